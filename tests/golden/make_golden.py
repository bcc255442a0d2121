"""Generate tests/golden/csgn_golden.json from the UNMODIFIED reference.

Run where /root/reference exists (after `make -C oracle`):

    python tests/golden/make_golden.py

Every value below is produced by the reference's public class API through
oracle/_ref/libcertfhe_ref.so (oracle/ref_shim.cpp): glibc srand(seed) after the
SecretKey is built, then SecretKey::encrypt, Ciphertext operator+ / operator*,
SecretKey::decrypt, Permutation(N), getInverse, operator+, applyPermutation.
The fixture travels to the GPU box; the reference does not.
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Ref, srand, words_per_block  # noqa: E402


def hexwords(a):
    return "".join("%016x" % int(x) for x in a)


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype="<u8").tobytes()).hexdigest()


def case(ref, N, D, seed, bits_a, bits_b, keep_full):
    rng = np.random.default_rng(seed)
    L = words_per_block(N)
    key = rng.permutation(N)[:D].astype(np.uint64)  # unsorted on purpose: order must not matter
    a = ref.encrypt_many(bits_a, N, D, key, seed=seed)
    b = ref.encrypt_many(bits_b, N, D, key, seed=seed + 1)
    prod, prod_bl = ref.mul(a, b, N, D)
    summ, sum_bl = ref.add(a, b, N, D)
    perm = ref.perm_generate(N, seed=seed + 2)
    inv = ref.perm_inverse(perm)
    ident = ref.perm_compose(perm, inv)
    pkey = ref.key_permute(N, D, key, perm)
    strict, strict_bl = ref.permute(a, N, D, perm)           # multi-block in, block 0 out
    per_block = np.concatenate([ref.permute(a[i * L:(i + 1) * L], N, D, perm)[0] for i in range(len(bits_a))])
    ct_size, sk_size = ref.sizes(a[:L], N, D, key)
    c = {
        "N": N, "D": D, "L": L, "seed": seed, "context": ref.context(N, D),
        "key": [int(x) for x in key],
        "bits_a": [int(x) for x in bits_a], "bits_b": [int(x) for x in bits_b],
        "enc_a": hexwords(a), "enc_b": hexwords(b),
        "dec_a": ref.decrypt(a, N, D, key), "dec_b": ref.decrypt(b, N, D, key),
        "mul_len": int(prod.size), "mul_sha256": digest(prod),
        "mul_bitlen_sha256": digest(prod_bl),
        "mul_head": hexwords(prod[: 2 * L]), "mul_tail": hexwords(prod[-L:]),
        "dec_mul": ref.decrypt(prod, N, D, key),
        "add_len": int(summ.size), "add_sha256": digest(summ), "add_bitlen_sha256": digest(sum_bl),
        "dec_add": ref.decrypt(summ, N, D, key),
        "perm_sha256": digest(perm), "perm_head": [int(x) for x in perm[:16]],
        "perm_inverse_sha256": digest(inv),
        "perm_compose_is_identity": bool(np.array_equal(ident, np.arange(N, dtype=np.uint64))),
        "permuted_key": [int(x) for x in pkey],
        "permute_strict_len": int(strict.size), "permute_strict": hexwords(strict),
        "permute_strict_bitlen": [int(x) for x in strict_bl] if L <= 32 else digest(strict_bl),
        "permute_each_block_sha256": digest(per_block),
        "dec_permuted": ref.decrypt(per_block, N, D, pkey),
        "ct_size_one_block": ct_size, "sk_size": sk_size,
    }
    if keep_full:
        c["mul"] = hexwords(prod)
        c["perm"] = [int(x) for x in perm]
        c["permute_each_block"] = hexwords(per_block)
    return c


def raw_case(ref, N, D, seed, T1, T2):
    """Raw (not encrypted) seeded blocks: what the bandwidth runs use."""
    rng = np.random.default_rng(seed)
    L = words_per_block(N)
    rem = N % 64
    pad = np.uint64(0xFFFFFFFFFFFFFFFF if rem == 0 else (0xFFFFFFFFFFFFFFFF << (64 - rem)) & 0xFFFFFFFFFFFFFFFF)
    a = rng.integers(0, 2**64, size=(T1, L), dtype=np.uint64)
    b = rng.integers(0, 2**64, size=(T2, L), dtype=np.uint64)
    a[:, L - 1] &= pad
    b[:, L - 1] &= pad
    a, b = a.reshape(-1), b.reshape(-1)
    key = rng.permutation(N)[:D].astype(np.uint64)
    prod, _ = ref.mul(a, b, N, D)
    return {"N": N, "D": D, "L": L, "seed": seed, "T1": T1, "T2": T2, "key": [int(x) for x in key],
            "rng": "numpy default_rng(seed): integers(0,2**64,(T1,L)) then (T2,L), last word & pad, permutation(N)[:D]",
            "a_sha256": digest(a), "b_sha256": digest(b),
            "mul_sha256": digest(prod), "dec_mul": ref.decrypt(prod, N, D, key),
            "dec_a": ref.decrypt(a, N, D, key)}


def main():
    ref = Ref()
    out = {"generator": "tests/golden/make_golden.py", "reference": "certfhe/CSGN src/*.cpp, -O3, public API",
           "cases": [], "raw_cases": []}
    out["cases"].append(case(ref, 1247, 16, 11, [1, 0, 1], [1, 1], keep_full=True))
    out["cases"].append(case(ref, 1247, 16, 12, [1] * 5, [1] * 3, keep_full=False))
    out["cases"].append(case(ref, 1247, 16, 13, [0, 1, 1, 0, 1, 1, 1], [1, 0, 1, 1], keep_full=False))
    out["cases"].append(case(ref, 65, 2, 21, [1, 1, 0, 1], [1, 1, 1], keep_full=True))
    out["cases"].append(case(ref, 191, 5, 22, [1, 0], [1], keep_full=True))      # L = 3 (odd)
    out["cases"].append(case(ref, 63, 4, 23, [1, 1], [0, 1, 1], keep_full=True))  # L = 1
    out["cases"].append(case(ref, 16383, 64, 31, [1, 0, 1], [1, 1], keep_full=False))
    out["raw_cases"].append(raw_case(ref, 1247, 16, 41, 37, 53))
    out["raw_cases"].append(raw_case(ref, 1247, 2, 42, 300, 200))
    out["raw_cases"].append(raw_case(ref, 16383, 3, 43, 9, 14))
    out["raw_cases"].append(raw_case(ref, 191, 1, 44, 33, 65))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csgn_golden.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
