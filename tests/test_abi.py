"""CPU: the C-ABI shared object loads, exports exactly what include/csgn.h declares,
and its host-only entry points behave.  No compute call is made without a GPU."""
import ctypes
import os
import subprocess

import pytest

from csgn_b200 import _native, build


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="module")
def lib():
    return _native.load()


def test_library_is_built_in_tree(lib):
    assert os.path.exists(build.libcsgn_path())
    assert os.path.dirname(build.libcsgn_path()).startswith(os.path.dirname(os.path.abspath(build.__file__)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    declared = _native.declared_symbols()
    assert len(declared) >= 35
    assert sorted(_native.SIGNATURES) == declared, "binding table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", build.libcsgn_path()], stdout=subprocess.PIPE, text=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert set(declared) <= exported


def test_kernels_are_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", build.libcsgn_path()], stdout=subprocess.PIPE, text=True).stdout
    archs = {tok for line in out.splitlines() for tok in line.replace(".", " ").split() if tok.startswith("sm_")}
    assert archs == {"sm_100a"}, archs


def test_words_per_block_matches_reference_context(lib):
    # src/Context.cpp:24-28
    for N, L in ((1247, 20), (16383, 256), (64, 1), (65, 2), (1, 1), (128, 2), (16384, 256)):
        assert lib.csgn_words_per_block(N) == L


def test_shard_range_partitions_contiguously(lib):
    for n in (0, 1, 7, 1000, 10**9 + 7):
        for world in (1, 2, 3, 8):
            nxt = 0
            for rank in range(world):
                first, count = ctypes.c_uint64(), ctypes.c_uint64()
                assert lib.csgn_shard_range(n, rank, world, ctypes.byref(first), ctypes.byref(count)) == 0
                assert first.value == nxt
                assert count.value in (n // world, n // world + 1)
                nxt += count.value
            assert nxt == n
    first, count = ctypes.c_uint64(), ctypes.c_uint64()
    assert lib.csgn_shard_range(10, 2, 2, ctypes.byref(first), ctypes.byref(count)) != 0
    assert b"rank" in lib.csgn_last_error()


def test_version_string(lib):
    assert b"sm_100a" in lib.csgn_version()


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback(lib):
    # run in a child so a half-initialised CUDA runtime cannot leak into this process
    code = (
        "import ctypes,sys; sys.path.insert(0, %r)\n"
        "from csgn_b200 import _native\n"
        "lib=_native.load(build_if_missing=False)\n"
        "rc=lib.csgn_init(0); print(rc, lib.csgn_last_error().decode())\n"
        "h=ctypes.c_void_p(); rc2=lib.csgn_buf_alloc(1,20,ctypes.byref(h)); print(rc2)\n"
        "sys.exit(0 if (rc==-4 and rc2==-1) else 1)\n" % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    r = subprocess.run(["python", "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout
    assert "no CPU path" in r.stdout


def test_header_is_plain_c99(tmp_path):
    """include/csgn.h is the drop-in boundary for C callers (cgo, JNI stubs, ctypes): it must compile as C, not only C++."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi_c.c"
    src.write_text('#include "csgn.h"\nint main(void) { return csgn_words_per_block(1247) == 20 ? 0 : 1; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-c", str(src),
                        "-o", str(tmp_path / "abi_c.o")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
