"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck python tools/sanitize_case.py

Touches every kernel (tiled and generic multiply; all decrypt forms incl. the lane-aligned kernel, the bulk-copy ring and
the fused publish + collect of a sharded decrypt at world size 1; concat/append; every form of the bit-sliced permute incl.
the bulk-copy prefetch, and the gather permute; batched encryption; the batch entry points; checksum; round 2: the fused
multiply->decrypt in every fold form, the rows fold, the window walk, the plane permute forms, deferred results, lazy sums,
batched uploads)
at sizes with ragged tails, and checks results against the oracle so that a silent corruption cannot pass.  compute-sanitizer is
closed on this pool; the script runs plain in the GPU suite (tests/test_gpu_parity.py) as an oracle-checked sweep."""
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
# ---- round-2 kernels: fused multiply->decrypt (lane-aligned and shared-memory fold), rows fold, plane permute, lazy sums
import ctypes  # noqa: E402
for N, D in ((1247, 2), (700, 2), (16383, 3), (191, 1), (33000, 2), (4097, 2), (8191, 2)):
    L = words_per_block(N)
    ctx = eng.Context(N, D)
    s = random_key(rng, N, D)
    key = eng.SecretKey(ctx, s)
    mask = np.zeros(L, dtype=np.uint64)
    for pos in s:
        mask[int(pos) >> 6] |= np.uint64(1 << (63 - (int(pos) & 63)))
    for T1, T2 in ((61, 97), (5, 130), (200, 1)) if L <= 64 else ((9, 14), (33, 5)):
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        a.reshape(-1, L)[::3] |= mask
        b.reshape(-1, L)[::2] |= mask
        ca, cb = eng.Ciphertext.from_host(a, ctx), eng.Ciphertext.from_host(b, ctx)
        want_words = o.mul(a, b, L)
        want = o.count_satisfied(want_words, N, s)
        for env in ({}, {"CSGN_MUL_ALIGN": "0"}, {"CSGN_MUL_ALIGN": "2", "CSGN_MUL_R": "3"}, {"CSGN_MUL_U": "1", "CSGN_MUL_R": "64"},
                    {"CSGN_MUL_TPB": "96", "CSGN_MUL_U": "4"}):
            os.environ.update(env)
            bit, cnt, prod = key.mul_decrypt(ca, cb, out="alloc")
            assert cnt == want and np.array_equal(prod.getValues(), want_words); n_checks += 1
            assert key.mul_decrypt(ca, cb) == (want & 1, want); n_checks += 1
            for k in env: del os.environ[k]
        assert key.count_satisfied(prod) == want; n_checks += 1                     # rows / wide / lanes / string fold by shape
        for bpi in ("1", "2", "4"):
            os.environ["CSGN_DEC_ROWS_BPI"] = bpi
            os.environ["CSGN_DEC_ROWS_MIN"] = "17"
            assert key.count_satisfied(prod) == want; n_checks += 1
        del os.environ["CSGN_DEC_ROWS_BPI"], os.environ["CSGN_DEC_ROWS_MIN"]
        perm = rng.permutation(N).astype(np.uint64)
        p = eng.Permutation(ctx, perm)
        wantp = o.permute_all(want_words, N, perm)
        for form in ("-1", "0", "6", "12", "17"):
            os.environ["CSGN_PERM_PLANE"] = form
            assert np.array_equal(prod.applyPermutation(p).getValues(), wantp); n_checks += 1
        del os.environ["CSGN_PERM_PLANE"]
        res = key.decrypt_deferred(prod)
        assert res.count() == want; n_checks += 1
# lazy sums: segments walked by decrypt / permute / left-operand product, flattened by everything else
N, D = 1247, 2
L = words_per_block(N)
ctx = eng.Context(N, D)
s = random_key(rng, N, D)
key = eng.SecretKey(ctx, s)
parts = [random_blocks(rng, t, N) for t in (7001, 6900)]
whole = np.concatenate(parts)
cts = [eng.Ciphertext.from_host(x, ctx) for x in parts]
rope = cts[0].add_lazy(cts[1])
assert rope.segments == 2 and key.count_satisfied(rope) == o.count_satisfied(whole, N, s); n_checks += 1
perm = rng.permutation(N).astype(np.uint64)
assert np.array_equal(rope.applyPermutation(eng.Permutation(ctx, perm)).getValues(), o.permute_all(whole, N, perm)); n_checks += 1
small = eng.Ciphertext.from_host(random_blocks(rng, 3, N), ctx)
assert np.array_equal((rope * small).getValues(), o.mul(whole, small.getValues(), L)); n_checks += 1
assert np.array_equal(rope.getValues(), whole) and rope.segments == 1; n_checks += 1
# batched upload: shared storage, views freed in any order
hosts = [np.ascontiguousarray(random_blocks(rng, t, N)) for t in (40, 1, 77)]
up = eng.UploadBatch([h.ctypes.data for h in hosts], [40, 1, 77], ctx)
ops = up.upload()
eng.sync()
for i in (2, 0, 1):
    view = eng.Ciphertext(ctypes.c_void_p(ops[i]), ctx)
    assert np.array_equal(view.getValues(), hosts[i]); n_checks += 1
    del view
eng.sync()
# the window walk of the decrypt fold: odd and even block lengths, whole and ragged runs, every form
for N, T in ((191, 1001), (1215, 333), (4097, 130), (4097, 2), (3197, 77), (12351, 41)):
    ctx = eng.Context(N, 3)
    s = random_key(rng, N, 3)
    key = eng.SecretKey(ctx, s)
    v = random_blocks(rng, T, N).reshape(T, -1)
    km = np.zeros(v.shape[1], dtype=np.uint64)
    for pos in s:
        km[int(pos) >> 6] |= np.uint64(1 << (63 - (int(pos) & 63)))
    v[rng.random(T) < 0.5] |= km
    v = np.ascontiguousarray(v.reshape(-1))
    ct = eng.Ciphertext.from_host(v, ctx)
    want = o.count_satisfied(v, N, s)
    for form in ("1", "2", "4", "6", "11"):
        os.environ["CSGN_DEC_WINDOW"] = form; os.environ["CSGN_DEC_WINDOW_ALL"] = "1"
        assert key.count_satisfied(ct) == want; n_checks += 1
    del os.environ["CSGN_DEC_WINDOW"], os.environ["CSGN_DEC_WINDOW_ALL"]
print("sanitize_case OK:", n_checks, "checks,", eng.launch_count(), "launches")
