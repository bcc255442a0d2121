"""SecretKey / Permutation files: host-only entry points of the C ABI -- no GPU, no csgn_init."""
import ctypes
import os

import numpy as np
import pytest

from csgn_b200 import _native

_vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def lib():
    return _native.load()


def _err(lib):
    return lib.csgn_last_error().decode()


def test_secret_key_file_round_trip(lib, tmp_path):
    N, D = 1247, 16
    pos = np.random.default_rng(1).permutation(N)[:D].astype(np.uint64)
    path = str(tmp_path / "k.sk").encode()
    assert lib.csgn_key_positions_save(path, N, D, pos.ctypes.data_as(_vp), pos.size) == 0, _err(lib)
    assert os.path.getsize(path) == 64 + 8 * D
    raw = np.fromfile(path, dtype=np.uint64)                 # the format, field by field (include/csgn.h)
    assert raw[:1].tobytes() == b"CSGNSK01" and [int(x) for x in raw[1:5]] == [N, D, 0, D]
    assert int(raw[5]) == int(np.bitwise_xor.reduce(pos)) and not raw[6:8].any() and np.array_equal(raw[8:], pos)
    n, d, cnt = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint64()
    assert lib.csgn_key_positions_load(path, ctypes.byref(n), ctypes.byref(d), None, 0, ctypes.byref(cnt)) == 0   # size query
    assert (n.value, d.value, cnt.value) == (N, D, D)
    got = np.zeros(D, dtype=np.uint64)
    assert lib.csgn_key_positions_load(path, ctypes.byref(n), ctypes.byref(d), got.ctypes.data_as(_vp), D, ctypes.byref(cnt)) == 0
    assert np.array_equal(got, pos)
    # too little room, a flipped bit, a truncated file, a foreign file: refused with a message
    assert lib.csgn_key_positions_load(path, None, None, got.ctypes.data_as(_vp), D - 1, ctypes.byref(cnt)) != 0
    assert "room for" in _err(lib)
    raw = bytearray(open(path, "rb").read())
    raw[64 + 3] ^= 0x10
    open(path, "wb").write(raw)
    assert lib.csgn_key_positions_load(path, None, None, got.ctypes.data_as(_vp), D, ctypes.byref(cnt)) != 0
    assert "checksum" in _err(lib) or "outside" in _err(lib)
    open(path, "wb").write(raw[:64 + 8 * D - 5])
    assert lib.csgn_key_positions_load(path, None, None, got.ctypes.data_as(_vp), D, ctypes.byref(cnt)) != 0
    assert "truncated" in _err(lib)
    # a position outside [0, N) never reaches a file
    bad = pos.copy()
    bad[2] = N
    assert lib.csgn_key_positions_save(path, N, D, bad.ctypes.data_as(_vp), bad.size) != 0


def test_permutation_file_round_trip(lib, tmp_path):
    n = 16383
    perm = np.random.default_rng(2).permutation(n).astype(np.uint64)
    path = str(tmp_path / "p.pm").encode()
    assert lib.csgn_perm_entries_save(path, perm.ctypes.data_as(_vp), n) == 0, _err(lib)
    raw = np.fromfile(path, dtype=np.uint64)
    assert raw[:1].tobytes() == b"CSGNPM01" and [int(x) for x in raw[1:5]] == [n, 0, 0, n] and np.array_equal(raw[8:], perm)
    cnt = ctypes.c_uint64()
    assert lib.csgn_perm_entries_load(path, None, 0, ctypes.byref(cnt)) == 0 and cnt.value == n
    got = np.zeros(n, dtype=np.uint64)
    assert lib.csgn_perm_entries_load(path, got.ctypes.data_as(_vp), n, ctypes.byref(cnt)) == 0
    assert np.array_equal(got, perm)
    # a key file is not a permutation file, and a file whose entries repeat is refused
    kpath = str(tmp_path / "k.sk").encode()
    pos = perm[:8].copy()
    assert lib.csgn_key_positions_save(kpath, n, 8, pos.ctypes.data_as(_vp), 8) == 0
    assert lib.csgn_perm_entries_load(kpath, got.ctypes.data_as(_vp), n, ctypes.byref(cnt)) != 0
    assert "not a CSGN permutation file" in _err(lib)
    twice = perm.copy()
    twice[5] = twice[6]
    assert lib.csgn_perm_entries_save(path, twice.ctypes.data_as(_vp), n) == 0      # saving does not validate bijectivity ...
    assert lib.csgn_perm_entries_load(path, got.ctypes.data_as(_vp), n, ctypes.byref(cnt)) != 0   # ... loading does
    assert "not a permutation" in _err(lib)
