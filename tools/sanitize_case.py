"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py

Touches every kernel (tiled and generic multiply; all decrypt forms incl. the lane-aligned kernel, the bulk-copy ring and
the fused publish + collect of a sharded decrypt at world size 1; concat/append; every form of the bit-sliced permute incl.
the bulk-copy prefetch, and the gather permute; batched encryption; the batch entry points; checksum) at sizes with ragged
tails, and checks results against the oracle so that a silent corruption cannot pass."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
from oracle.pyoracle import Oracle, random_blocks, random_key, words_per_block

eng.init(0)
o = Oracle()
rng = np.random.default_rng(5)
n_checks = 0
for N in (1247, 16383, 191, 2048):
    L = words_per_block(N)
    ctx = eng.Context(N, 2)
    shapes = [(37, 53), (5, 1), (1, 9), (300, 7), (3, 65)] if L <= 64 else [(9, 14), (40, 3)]
    for T1, T2 in shapes:
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        ca, cb = eng.Ciphertext.from_host(a, ctx), eng.Ciphertext.from_host(b, ctx)
        for env in ({}, {"CSGN_MUL_U": "1"}, {"CSGN_MUL_U": "4", "CSGN_MUL_R": "2"}, {"CSGN_MUL_GENERIC": "1"}):
            os.environ.update(env)
            assert np.array_equal((ca * cb).getValues(), o.mul(a, b, L)); n_checks += 1
            for k in env: del os.environ[k]
        s = random_key(rng, N, 2)
        key = eng.SecretKey(ctx, s)
        prod = ca * cb
        want = o.count_satisfied(o.mul(a, b, L), N, s)
        variants = ("0", "1", "2", "5", "6", "8", "12", "13") if N == 1247 else ("0",)
        for var in variants:
            os.environ["CSGN_DEC_VARIANT"] = var
            assert key.count_satisfied(prod) == want; n_checks += 1
        del os.environ["CSGN_DEC_VARIANT"]
        os.environ["CSGN_DEC_GENERIC"] = "1"
        assert key.count_satisfied(prod) == want; n_checks += 1
        del os.environ["CSGN_DEC_GENERIC"]
        cat = ca + cb
        cat += ca
        assert np.array_equal(cat.getValues(), o.concat(o.concat(a, b), a)); n_checks += 1
        perm = rng.permutation(N).astype(np.uint64)
        p = eng.Permutation(ctx, perm)
        assert np.array_equal(cat.applyPermutation(p).getValues(), o.permute_all(cat.getValues(), N, perm)); n_checks += 1
        if N in (1247, 16383):
            for var in ("1", "2", "3", "4", "5", "8"):
                os.environ["CSGN_PERM_VARIANT"] = var
                assert np.array_equal(cat.applyPermutation(p).getValues(), o.permute_all(cat.getValues(), N, perm)); n_checks += 1
            del os.environ["CSGN_PERM_VARIANT"]
        os.environ["CSGN_PERM_GATHER"] = "1"
        assert np.array_equal(ca.applyPermutation(p).getValues(), o.permute_all(a, N, perm)); n_checks += 1
        del os.environ["CSGN_PERM_GATHER"]
        assert prod.checksum() == o.checksum(prod.getValues()); n_checks += 1
        # batch entry points (lanes) and the sharded fold at world size 1 (publish + collect in the kernel)
        prods = eng.mul_batch([ca, cb, ca], [cb, ca, ca])
        for got, (x, y) in zip(prods, ((a, b), (b, a), (a, a))):
            assert np.array_equal(got.getValues(), o.mul(x, y, L)); n_checks += 1
        bits, cnts = key.decrypt_batch(prods)
        assert cnts == [o.count_satisfied(g.getValues(), N, s) for g in prods]; n_checks += 1
        comm = eng.PeerComm(0, 1)
        assert comm.decrypt(key, prod) == (want & 1, want); n_checks += 1
        del comm
        plain = rng.integers(0, 2, size=41).astype(np.uint8)
        fresh = key.encrypt_batch(plain, seed=7)
        assert np.array_equal(fresh.getValues(), o.encrypt_batch(plain, N, s, 7)); n_checks += 1
eng.sync()
print("sanitize_case OK:", n_checks, "checks,", eng.launch_count(), "launches")
