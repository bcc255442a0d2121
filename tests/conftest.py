import json
import os
import sys

import numpy as np
import pytest

os.environ.setdefault("CSGN_TUNING", "1")   # the launchers honour CSGN_* knobs only when this is set at csgn_init

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.pyoracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference build; present wherever oracle/_ref was built or shipped."""
    from oracle import pyoracle
    if not pyoracle.ref_available():
        pytest.skip("oracle/_ref/libcertfhe_ref.so not built (needs /root/reference at build time)")
    return pyoracle.Ref()


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "csgn_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def engine():
    """The CUDA path through the C ABI.  Fails (never skips) if the library or GPU is missing."""
    from csgn_b200 import engine as eng
    eng.init(-1)
    return eng


def unhex(s):
    return np.array([int(s[i:i + 16], 16) for i in range(0, len(s), 16)], dtype=np.uint64)


def sha(a):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()
