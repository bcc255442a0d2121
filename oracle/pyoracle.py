"""ctypes bindings for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Oracle``  -> oracle/libcsgn_oracle.so, the plain-C restatement (csgn_oracle.c).
* ``Ref``     -> oracle/_ref/libcertfhe_ref.so, the unmodified reference behind
                 oracle/ref_shim.cpp (present when oracle/Makefile could see
                 /root/reference at build time; the built file travels to the GPU box).

Both operate on numpy uint64 arrays.  glibc ``srand`` is exposed so that callers can
seed the shared ``rand()`` stream the reference draws from.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u64 = ctypes.c_uint64
_libc = ctypes.CDLL(None)
_libc.srand.argtypes = [ctypes.c_uint]
_libc.rand.restype = ctypes.c_int


def srand(seed):
    _libc.srand(seed)


def crand():
    return _libc.rand()


def _p(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def build_oracle():
    """Compile the C restatement (gcc, <1 s).  Building the checker is not using it."""
    subprocess.run(["make", "-s", "-C", _HERE, "libcsgn_oracle.so"], check=True)


def words_per_block(N):
    return N // 64 + (1 if N % 64 else 0)


def pad_mask(N):
    rem = N % 64
    return np.uint64(0xFFFFFFFFFFFFFFFF if rem == 0 else (0xFFFFFFFFFFFFFFFF << (64 - rem)) & 0xFFFFFFFFFFFFFFFF)


def random_blocks(rng, T, N):
    """T raw blocks of seeded random words with the pad bits of the last word zeroed."""
    L = words_per_block(N)
    w = rng.integers(0, 2**64, size=(T, L), dtype=np.uint64)
    w[:, L - 1] &= pad_mask(N)
    return w.reshape(-1)


def random_key(rng, N, D):
    return np.sort(rng.choice(N, size=D, replace=False)).astype(np.uint64)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(_HERE, "libcsgn_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        self.lib = lib = ctypes.CDLL(path)
        lib.csgn_oracle_words_per_block.restype = _u64
        lib.csgn_oracle_words_per_block.argtypes = [_u64]
        lib.csgn_oracle_S.restype = _u64
        lib.csgn_oracle_S.argtypes = [_u64, _u64]
        lib.csgn_oracle_canonical_bitlen.argtypes = [_u64, _u64, _u64p]
        lib.csgn_oracle_mul.argtypes = [_u64p, _u64, _u64p, _u64, _u64, _u64p]
        lib.csgn_oracle_mul_checksum.argtypes = [_u64p, _u64, _u64p, _u64, _u64, _u64p, _u64p, _u64p]
        lib.csgn_oracle_checksum.argtypes = [_u64p, _u64, _u64p, _u64p, _u64p]
        lib.csgn_oracle_concat.argtypes = [_u64p, _u64, _u64p, _u64, _u64p]
        lib.csgn_oracle_key_mask.argtypes = [_u64, _u64p, _u64, _u64p]
        for f in (lib.csgn_oracle_decrypt, lib.csgn_oracle_decrypt_unpacked, lib.csgn_oracle_count_satisfied):
            f.restype = _u64
            f.argtypes = [_u64p, _u64, _u64, _u64p, _u64]
        lib.csgn_oracle_permute_block.argtypes = [_u64p, _u64, _u64p, _u64p]
        lib.csgn_oracle_permute_all.argtypes = [_u64p, _u64, _u64, _u64p, _u64p]
        lib.csgn_oracle_key_permute.restype = _u64
        lib.csgn_oracle_key_permute.argtypes = [_u64, _u64p, _u64, _u64p, _u64p]
        lib.csgn_oracle_perm_inverse.argtypes = [_u64p, _u64, _u64p]
        lib.csgn_oracle_perm_compose.argtypes = [_u64p, _u64p, _u64, _u64p]
        lib.csgn_oracle_encrypt.argtypes = [ctypes.c_int, _u64, _u64, _u64p, _u64p]
        lib.csgn_oracle_perm_generate.argtypes = [_u64, _u64p]
        lib.csgn_oracle_keygen.argtypes = [_u64, _u64, _u64p]
        lib.csgn_oracle_bits_text.argtypes = [_u64p, _u64, _u64, ctypes.c_char_p]
        lib.csgn_oracle_encrypt_batch.argtypes = [ctypes.POINTER(ctypes.c_uint8), _u64, _u64, _u64, _u64p, _u64, _u64, _u64p]

    def canonical_bitlen(self, N, T):
        out = np.empty(T * words_per_block(N), dtype=np.uint64)
        self.lib.csgn_oracle_canonical_bitlen(N, T, _p(out))
        return out

    def mul(self, a, b, L):
        a, b = _arr(a), _arr(b)
        T1, T2 = a.size // L, b.size // L
        out = np.empty(T1 * T2 * L, dtype=np.uint64)
        self.lib.csgn_oracle_mul(_p(a), T1, _p(b), T2, L, _p(out))
        return out

    def mul_checksum(self, a, b, L):
        a, b = _arr(a), _arr(b)
        x, s, h = _u64(), _u64(), _u64()
        self.lib.csgn_oracle_mul_checksum(_p(a), a.size // L, _p(b), b.size // L, L,
                                          ctypes.byref(x), ctypes.byref(s), ctypes.byref(h))
        return x.value, s.value, h.value

    def checksum(self, v):
        v = _arr(v)
        x, s, h = _u64(), _u64(), _u64()
        self.lib.csgn_oracle_checksum(_p(v), v.size, ctypes.byref(x), ctypes.byref(s), ctypes.byref(h))
        return x.value, s.value, h.value

    def concat(self, a, b):
        a, b = _arr(a), _arr(b)
        out = np.empty(a.size + b.size, dtype=np.uint64)
        self.lib.csgn_oracle_concat(_p(a), a.size, _p(b), b.size, _p(out))
        return out

    def key_mask(self, N, s):
        s = _arr(s)
        out = np.empty(words_per_block(N), dtype=np.uint64)
        self.lib.csgn_oracle_key_mask(N, _p(s), s.size, _p(out))
        return out

    def decrypt(self, v, N, s, unpacked=False):
        v, s = _arr(v), _arr(s)
        T = v.size // words_per_block(N)
        f = self.lib.csgn_oracle_decrypt_unpacked if unpacked else self.lib.csgn_oracle_decrypt
        return int(f(_p(v), T, N, _p(s), s.size))

    def count_satisfied(self, v, N, s):
        v, s = _arr(v), _arr(s)
        return int(self.lib.csgn_oracle_count_satisfied(_p(v), v.size // words_per_block(N), N, _p(s), s.size))

    def permute_block(self, blk, N, perm):
        blk, perm = _arr(blk), _arr(perm)
        out = np.empty(words_per_block(N), dtype=np.uint64)
        self.lib.csgn_oracle_permute_block(_p(blk), N, _p(perm), _p(out))
        return out

    def permute_all(self, v, N, perm):
        v, perm = _arr(v), _arr(perm)
        out = np.empty_like(v)
        self.lib.csgn_oracle_permute_all(_p(v), v.size // words_per_block(N), N, _p(perm), _p(out))
        return out

    def key_permute(self, N, s, perm):
        s, perm = _arr(s), _arr(perm)
        out = np.empty(s.size, dtype=np.uint64)
        n = self.lib.csgn_oracle_key_permute(N, _p(s), s.size, _p(perm), _p(out))
        return out[:n]

    def perm_inverse(self, perm):
        perm = _arr(perm)
        out = np.empty_like(perm)
        self.lib.csgn_oracle_perm_inverse(_p(perm), perm.size, _p(out))
        return out

    def perm_compose(self, p, q):
        p, q = _arr(p), _arr(q)
        out = np.empty_like(p)
        self.lib.csgn_oracle_perm_compose(_p(p), _p(q), p.size, _p(out))
        return out

    def encrypt(self, bit, N, D, s):
        s = _arr(s)
        out = np.empty(words_per_block(N), dtype=np.uint64)
        self.lib.csgn_oracle_encrypt(int(bit), N, D, _p(s), _p(out))
        return out

    def perm_generate(self, n):
        out = np.empty(n, dtype=np.uint64)
        self.lib.csgn_oracle_perm_generate(n, _p(out))
        return out

    def keygen(self, N, D):
        out = np.empty(D, dtype=np.uint64)
        self.lib.csgn_oracle_keygen(N, D, _p(out))
        return out

    def encrypt_batch(self, bits, N, s, seed, first_block=0):
        bits = np.ascontiguousarray(np.asarray(bits, dtype=np.uint8))
        s = _arr(s)
        out = np.empty(bits.size * words_per_block(N), dtype=np.uint64)
        self.lib.csgn_oracle_encrypt_batch(bits.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), bits.size, first_block,
                                           N, _p(s), s.size, seed, _p(out))
        return out

    def bits_text(self, v, N):
        v = _arr(v)
        T = v.size // words_per_block(N)
        buf = ctypes.create_string_buffer(T * N + 1)
        self.lib.csgn_oracle_bits_text(_p(v), T, N, buf)
        return buf.value.decode()


def ref_path(opt="O3"):
    name = "libcertfhe_ref.so" if opt == "O3" else "libcertfhe_ref_O0.so"
    return os.path.join(_HERE, "_ref", name)


def ref_available(opt="O3"):
    return os.path.exists(ref_path(opt))


class Ref:
    """The unmodified reference, through its public class API (oracle/ref_shim.cpp)."""

    def __init__(self, opt="O3"):
        path = ref_path(opt)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        self.lib = lib = ctypes.CDLL(path)
        vp = ctypes.c_void_p
        lib.ref_ct_new.restype = vp
        lib.ref_ct_new.argtypes = [_u64p, _u64, _u64, _u64]
        lib.ref_ct_free.argtypes = [vp]
        lib.ref_ct_len.restype = _u64
        lib.ref_ct_len.argtypes = [vp]
        lib.ref_ct_words.argtypes = [vp, _u64p]
        lib.ref_ct_bitlen.argtypes = [vp, _u64p]
        lib.ref_ct_size.restype = ctypes.c_long
        lib.ref_ct_size.argtypes = [vp]
        for f in (lib.ref_ct_mul, lib.ref_ct_add):
            f.restype = vp
            f.argtypes = [vp, vp]
        for f in (lib.ref_ct_mul_inplace, lib.ref_ct_add_inplace):
            f.argtypes = [vp, vp]
        lib.ref_ct_permute.restype = vp
        lib.ref_ct_permute.argtypes = [vp, _u64p, _u64]
        lib.ref_sk_new.restype = vp
        lib.ref_sk_new.argtypes = [_u64, _u64, _u64p]
        lib.ref_sk_free.argtypes = [vp]
        lib.ref_sk_size.restype = ctypes.c_long
        lib.ref_sk_size.argtypes = [vp]
        lib.ref_sk_decrypt.restype = ctypes.c_int
        lib.ref_sk_decrypt.argtypes = [vp, vp]
        lib.ref_sk_encrypt.restype = vp
        lib.ref_sk_encrypt.argtypes = [vp, ctypes.c_int]
        lib.ref_sk_permute.argtypes = [vp, _u64p, _u64, _u64p]
        lib.ref_perm_generate.argtypes = [_u64, _u64p]
        lib.ref_perm_inverse.argtypes = [_u64p, _u64, _u64p]
        lib.ref_perm_compose.restype = _u64
        lib.ref_perm_compose.argtypes = [_u64p, _u64, _u64p, _u64, _u64p]
        lib.ref_context.argtypes = [_u64, _u64, _u64p]
        lib.ref_bench_mul_decrypt.restype = ctypes.c_int
        lib.ref_bench_mul_decrypt.argtypes = [_u64, _u64, _u64, _u64, ctypes.c_int, ctypes.c_int, _u64,
                                              ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), _u64p]

    # -- handles ---------------------------------------------------------
    def ct(self, words, N, D):
        words = _arr(words)
        return self.lib.ref_ct_new(_p(words), words.size, N, D)

    def ct_free(self, h):
        self.lib.ref_ct_free(h)

    def ct_words(self, h):
        out = np.empty(self.lib.ref_ct_len(h), dtype=np.uint64)
        if out.size:
            self.lib.ref_ct_words(h, _p(out))
        return out

    def ct_bitlen(self, h):
        out = np.empty(self.lib.ref_ct_len(h), dtype=np.uint64)
        if out.size:
            self.lib.ref_ct_bitlen(h, _p(out))
        return out

    def sk(self, N, D, s):
        s = _arr(s)
        return self.lib.ref_sk_new(N, D, _p(s))

    # -- value-level conveniences (build handles, run, free) --------------
    def _binary(self, fn, a, b, N, D):
        ha, hb = self.ct(a, N, D), self.ct(b, N, D)
        hc = fn(ha, hb)
        out, bl = self.ct_words(hc), self.ct_bitlen(hc)
        for h in (ha, hb, hc):
            self.ct_free(h)
        return out, bl

    def mul(self, a, b, N, D):
        return self._binary(self.lib.ref_ct_mul, a, b, N, D)

    def add(self, a, b, N, D):
        return self._binary(self.lib.ref_ct_add, a, b, N, D)

    def _inplace(self, fn, a, b, N, D):
        ha, hb = self.ct(a, N, D), self.ct(b, N, D)
        fn(ha, hb)
        out, bl = self.ct_words(ha), self.ct_bitlen(ha)
        self.ct_free(ha)
        self.ct_free(hb)
        return out, bl

    def mul_inplace(self, a, b, N, D):
        return self._inplace(self.lib.ref_ct_mul_inplace, a, b, N, D)

    def add_inplace(self, a, b, N, D):
        return self._inplace(self.lib.ref_ct_add_inplace, a, b, N, D)

    def decrypt(self, v, N, D, s):
        hk, hc = self.sk(N, D, s), self.ct(v, N, D)
        bit = self.lib.ref_sk_decrypt(hk, hc)
        self.lib.ref_sk_free(hk)
        self.ct_free(hc)
        return int(bit)

    def encrypt(self, bit, N, D, s, seed=None):
        """Fresh encryption.  The SecretKey ctor reseeds rand(); `seed` is applied after it."""
        hk = self.sk(N, D, s)
        if seed is not None:
            srand(seed)
        hc = self.lib.ref_sk_encrypt(hk, int(bit))
        out = self.ct_words(hc)
        self.ct_free(hc)
        self.lib.ref_sk_free(hk)
        return out

    def encrypt_many(self, bits, N, D, s, seed):
        hk = self.sk(N, D, s)
        srand(seed)
        outs = []
        for b in bits:
            hc = self.lib.ref_sk_encrypt(hk, int(b))
            outs.append(self.ct_words(hc))
            self.ct_free(hc)
        self.lib.ref_sk_free(hk)
        return np.concatenate(outs) if outs else np.empty(0, dtype=np.uint64)

    def permute(self, v, N, D, perm):
        perm = _arr(perm)
        hc = self.ct(v, N, D)
        hp = self.lib.ref_ct_permute(hc, _p(perm), perm.size)
        out, bl = self.ct_words(hp), self.ct_bitlen(hp)
        self.ct_free(hc)
        self.ct_free(hp)
        return out, bl

    def key_permute(self, N, D, s, perm):
        perm = _arr(perm)
        hk = self.sk(N, D, s)
        out = np.empty(D, dtype=np.uint64)
        self.lib.ref_sk_permute(hk, _p(perm), perm.size, _p(out))
        self.lib.ref_sk_free(hk)
        return out

    def perm_generate(self, n, seed=None):
        if seed is not None:
            srand(seed)
        out = np.empty(n, dtype=np.uint64)
        self.lib.ref_perm_generate(n, _p(out))
        return out

    def perm_inverse(self, perm):
        perm = _arr(perm)
        out = np.empty_like(perm)
        self.lib.ref_perm_inverse(_p(perm), perm.size, _p(out))
        return out

    def perm_compose(self, a, b):
        a, b = _arr(a), _arr(b)
        out = np.empty(max(a.size, 1), dtype=np.uint64)
        n = self.lib.ref_perm_compose(_p(a), a.size, _p(b), b.size, _p(out))
        return out[:n]

    def context(self, N, D):
        out = np.empty(4, dtype=np.uint64)
        self.lib.ref_context(N, D, _p(out))
        return [int(x) for x in out]

    def sizes(self, v, N, D, s):
        hk, hc = self.sk(N, D, s), self.ct(v, N, D)
        r = (int(self.lib.ref_ct_size(hc)), int(self.lib.ref_sk_size(hk)))
        self.ct_free(hc)
        self.lib.ref_sk_free(hk)
        return r

    def bench_mul_decrypt(self, N, D, T1, T2, threads=1, reps=1, seed=1):
        m, d, par = ctypes.c_double(), ctypes.c_double(), _u64()
        rc = self.lib.ref_bench_mul_decrypt(N, D, T1, T2, threads, reps, seed,
                                            ctypes.byref(m), ctypes.byref(d), ctypes.byref(par))
        if rc:
            raise RuntimeError("ref_bench_mul_decrypt failed")
        return m.value, d.value, par.value
