"""csgn_b200 -- B200-native evaluation engine for CSGN / certFHE ciphertexts.

Layout (only what the hot path needs):
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/csgn.h)
  certfhe/   the certFHE C++ classes (drop-in for the reference's src/*.h) over the C ABI
  engine.py  the same boundary for Python callers (tests, bench.py, torchrun launches)
  build.py   in-tree nvcc / g++ build

There is no CPU compute path: importing works anywhere, running needs a B200.
"""
from . import build  # noqa: F401

__all__ = ["build", "engine"]
__version__ = "0.1"
