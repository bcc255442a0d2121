"""GPU: round-2 kernels and scheduling through the C ABI, bit-exact against the oracle.

  * the fused multiply -> decrypt kernel (product words AND the satisfied-block count from one launch; count-only
    form), tiled and flat forms, every unit width (16-byte and 8-byte units: even and odd L), ragged tiles;
  * the flat ("chain shape") multiply;
  * the decrypt fold for every block shape (lane-aligned for 1..16 units, wide for multiples of 32, fail-string);
  * automatic lanes / deferred results / uploads shared by several streams (cross-stream ordering).
"""
import os

import numpy as np
import pytest

from oracle.pyoracle import random_blocks, random_key, words_per_block

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kv):
        self.kv = {k: str(v) for k, v in kv.items()}

    def __enter__(self):
        self.old = {k: os.environ.get(k) for k in self.kv}
        os.environ.update(self.kv)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def key_mask(N, s):
    L = words_per_block(N)
    m = np.zeros(L, dtype=np.uint64)
    for p in s:
        m[int(p) >> 6] |= np.uint64(1 << (63 - (int(p) & 63)))
    return m


def planted(rng, T, N, s, frac=0.3):
    """random blocks, a fraction of them with every key bit set (so that products have satisfied blocks)"""
    L = words_per_block(N)
    w = random_blocks(rng, T, N).reshape(T, L)
    rows = np.nonzero(rng.random(T) < frac)[0]
    w[rows] |= key_mask(N, s)
    return w.reshape(-1)


def check_fused(engine, oracle, N, D, T1, T2, rng, tag=""):
    L = words_per_block(N)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    a, b = planted(rng, T1, N, s), planted(rng, T2, N, s)
    ca, cb = engine.Ciphertext.from_host(a, ctx), engine.Ciphertext.from_host(b, ctx)
    want_words = oracle.mul(a, b, L)
    want = oracle.count_satisfied(want_words, N, s)
    bit, count, prod = key.mul_decrypt(ca, cb, out="alloc")
    assert count == want and bit == (want & 1), (tag, N, D, T1, T2, count, want)
    assert np.array_equal(prod.getValues(), want_words), (tag, N, T1, T2)
    bit2, count2 = key.mul_decrypt(ca, cb)                     # count only: nothing is stored
    assert (bit2, count2) == (bit, count), (tag, N, D, T1, T2, "count-only")
    into = engine.Ciphertext.empty(T1 * T2, ctx)
    bit3, count3 = key.mul_decrypt(ca, cb, out=into)
    assert (bit3, count3) == (bit, count) and np.array_equal(into.getValues(), want_words)
    return want


FUSED_SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (37, 53), (300, 200), (1000, 3), (3, 1000), (129, 65), (64, 640)]


@pytest.mark.parametrize("N,D", [(1247, 2), (1247, 16), (16383, 3), (16383, 64), (65, 1), (191, 2), (63, 2), (2048, 3),
                                 (4097, 2), (8191, 2), (33000, 2), (70000, 1), (129, 1), (640, 2)])
def test_fused_mul_decrypt_matches_oracle(engine, oracle, N, D):
    rng = np.random.default_rng(N * 11 + D)
    L = words_per_block(N)
    shapes = FUSED_SHAPES if L <= 64 else [(1, 1), (1, 5), (5, 1), (9, 14), (40, 33), (70, 3)]
    seen = 0
    for T1, T2 in shapes:
        seen += check_fused(engine, oracle, N, D, T1, T2, rng)
    assert seen > 0          # the planted blocks made the fold count something


@pytest.mark.parametrize("knobs", [dict(CSGN_MUL_U=1), dict(CSGN_MUL_U=2, CSGN_MUL_R=1), dict(CSGN_MUL_U=8, CSGN_MUL_R=5),
                                   dict(CSGN_MUL_U=4, CSGN_MUL_R=64, CSGN_MUL_GRID=3), dict(CSGN_MUL_TPB=160),
                                   dict(CSGN_MUL_TPB=512, CSGN_MUL_U=8), dict(CSGN_MUL_U=1, CSGN_MUL_R=64),
                                   dict(CSGN_MUL_FLAT=1), dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=1),
                                   dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=8, CSGN_MUL_GRID=5),
                                   dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=2, CSGN_MUL_TPB=64),
                                   dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_CTAS_PER_SM=1),
                                   # the shared-memory fold instead of the lane-aligned one, and the lane-aligned one at
                                   # every unroll, tiny and maximal row counts, odd CTA sizes, a 3-CTA grid
                                   dict(CSGN_MUL_ALIGN=0), dict(CSGN_MUL_ALIGN=0, CSGN_MUL_U=4, CSGN_MUL_R=16),
                                   dict(CSGN_MUL_ALIGN=1, CSGN_MUL_U=1, CSGN_MUL_R=1), dict(CSGN_MUL_ALIGN=1, CSGN_MUL_U=2, CSGN_MUL_R=32),
                                   dict(CSGN_MUL_ALIGN=1, CSGN_MUL_U=4, CSGN_MUL_R=16, CSGN_MUL_GRID=3),
                                   dict(CSGN_MUL_ALIGN=1, CSGN_MUL_U=8, CSGN_MUL_R=8, CSGN_MUL_TPB=288),
                                   dict(CSGN_MUL_ALIGN=1, CSGN_MUL_U=1, CSGN_MUL_R=64, CSGN_MUL_TPB=64)])
def test_fused_every_kernel_form(engine, oracle, knobs):
    rng = np.random.default_rng(77)
    for N, D in ((1247, 2), (16383, 3), (191, 1), (700, 2), (1000, 3), (330, 1)):      # 11 / 16 (8 units) / 6 words per block
        for T1, T2 in ((61, 97), (5, 700), (200, 1), (2000, 13), (700, 41)):
            if N == 16383:
                T1, T2 = max(1, T1 // 4), max(1, T2 // 4)
            with _Env(**knobs):
                check_fused(engine, oracle, N, D, T1, T2, rng, tag=str(knobs))


@pytest.mark.parametrize("knobs", [dict(CSGN_MUL_FLAT=1), dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=1),
                                   dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=2, CSGN_MUL_GRID=7),
                                   dict(CSGN_MUL_FLAT=1, CSGN_MUL_FLAT_U=8), dict(CSGN_MUL_FLAT=1, CSGN_MUL_TPB=128)])
def test_flat_multiply_matches_oracle(engine, oracle, knobs):
    """The chain-shape kernel (whole right operand in shared memory, output walked as one flat stream) -- a losing
    variant, compiled only with -DCSGN_BUILD_VARIANTS."""
    if not engine.has_variants():
        pytest.skip("libcsgn.so was built without the tuning variants")
    rng = np.random.default_rng(5)
    for N in (1247, 16383, 191, 2048):
        L = words_per_block(N)
        ctx = engine.Context(N, 4)
        for T1, T2 in ((2, 7), (100, 125), (3000, 25), (777, 13), (50, 200), (4001, 7)):
            if L > 64:
                T1, T2 = max(2, T1 // 8), max(1, T2 // 4)
            a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
            with _Env(**knobs):
                got = (engine.Ciphertext.from_host(a, ctx) * engine.Ciphertext.from_host(b, ctx)).getValues()
            assert np.array_equal(got, oracle.mul(a, b, L)), (knobs, N, T1, T2)


def test_fused_batch_and_views(engine, oracle):
    """csgn_mul_count_batch_async over the library's lanes: products into caller-owned views, counts on the device."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(8)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    P, T1, T2 = 6, 90, 70
    hosts = [(planted(rng, T1, N, s), planted(rng, T2, N, s)) for _ in range(P)]
    ca = [engine.Ciphertext.from_host(h[0], ctx) for h in hosts]
    cb = [engine.Ciphertext.from_host(h[1], ctx) for h in hosts]
    dev = torch.device("cuda", torch.cuda.current_device())
    out = torch.zeros((P, T1 * T2 * L), dtype=torch.int64, device=dev)
    vo = [engine.Ciphertext.from_tensor(out[p], ctx) for p in range(P)]
    counts = torch.zeros(P, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    engine.mul_count_batch_async(key, ca, cb, counts.data_ptr(), out=vo)
    engine.sync()
    torch.cuda.synchronize()
    got_counts = counts.cpu().numpy()
    for p in range(P):
        want_words = oracle.mul(hosts[p][0], hosts[p][1], L)
        assert np.array_equal(out[p].cpu().numpy().view(np.uint64), want_words), p
        assert got_counts[p] == oracle.count_satisfied(want_words, N, s), p
    # count only, and library-allocated products
    counts.zero_()
    torch.cuda.synchronize()
    engine.mul_count_batch_async(key, ca, cb, counts.data_ptr())
    engine.sync()
    assert np.array_equal(counts.cpu().numpy(), got_counts)
    prods = engine.mul_count_batch_async(key, ca, cb, counts.data_ptr(), out="alloc")
    engine.sync()
    for p in range(P):
        assert np.array_equal(prods[p].getValues(), oracle.mul(hosts[p][0], hosts[p][1], L))


def test_fused_sharded_world1(engine, oracle):
    """multiply + fold + publish + collect in one kernel (world size 1: own mailbox only)."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(18)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    comm = engine.PeerComm(0, 1)
    dev = torch.device("cuda", torch.cuda.current_device())
    totals = torch.zeros(4, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    wants = []
    cts = []
    for i in range(4):
        a, b = planted(rng, 50 + i, N, s), planted(rng, 40, N, s)
        ca, cb = engine.Ciphertext.from_host(a, ctx), engine.Ciphertext.from_host(b, ctx)
        cts.append((ca, cb))
        wants.append(oracle.count_satisfied(oracle.mul(a, b, L), N, s))
    out = engine.Ciphertext.empty(53 * 40, ctx)
    for i, (ca, cb) in enumerate(cts):
        last = i == 3
        comm.mul_push(key, ca, cb, out=out if last else None, collect_n=4 if last else 0,
                      device_totals_ptr=totals.data_ptr() if last else 0)
    engine.sync()
    assert totals.cpu().tolist() == wants
    assert oracle.count_satisfied(out.getValues(), N, s) == wants[3]


# ---------------------------------------------------------------------------
# the decrypt fold for every block shape
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("L", list(range(1, 35)) + [48, 64, 65, 96, 128, 160, 192, 224, 256, 320, 512, 513, 1024, 1025, 1100])
def test_decrypt_every_block_shape(engine, oracle, L):
    """N = 64*L - 3: L words per block.  Even L up to 32 -> lanes (16-byte units); odd L up to 16 -> lanes (8-byte
    units); 16-byte unit counts that are multiples of 32 -> wide; the rest -> fail-string (16- or 8-byte units) or the
    warp-per-block kernel beyond 512 units."""
    N, D = 64 * L - 3, 3
    rng = np.random.default_rng(L)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    for T in ([0, 1, 2, 31, 33, 257, 1000, 4099] if L <= 64 else [1, 33, 130]):
        v = planted(rng, T, N, s, 0.4) if T else random_blocks(rng, 0, N)
        ct = engine.Ciphertext.from_host(v, ctx)
        want = oracle.count_satisfied(v, N, s)
        assert key.count_satisfied(ct) == want, (L, T)
        with _Env(CSGN_DEC_STRING=1):
            assert key.count_satisfied(ct) == want, (L, T, "string")
        with _Env(CSGN_DEC_GENERIC=1):
            assert key.count_satisfied(ct) == want, (L, T, "generic")
        if 3 <= L <= 999:                   # the window walk (default for odd L = 3, >= 17 and some even L): every form, any L
            for form in (2, 3, 4, 6, 11):
                with _Env(CSGN_DEC_WINDOW=form, CSGN_DEC_WINDOW_ALL=1):
                    assert key.count_satisfied(ct) == want, (L, T, "window form", form)
            with _Env(CSGN_DEC_WINDOW=0):
                assert key.count_satisfied(ct) == want, (L, T, "without the window walk")
        if L % 2 and L >= 17:               # odd L: the double-block kernel, every blocks-per-iteration form; and without it
            for bpi in (1, 2, 4):
                with _Env(CSGN_DEC_ROWS_BPI=bpi, CSGN_DEC_PAIRS_MIN=17):
                    assert key.count_satisfied(ct) == want, (L, T, "pairs", bpi)
            with _Env(CSGN_DEC_PAIRS_MIN=100000):
                assert key.count_satisfied(ct) == want, (L, T, "string over double blocks")
            with _Env(CSGN_DEC_PAIRS_MIN=17):
                assert key.count_satisfied(ct) == want, (L, T, "rows over double blocks")
            with _Env(CSGN_DEC_PAIRS_MIN=100000, CSGN_DEC_STRING_PAIRS=0):
                assert key.count_satisfied(ct) == want, (L, T, "8-byte units")
        if T > 3:
            # a 16-byte-misaligned view of an even-L ciphertext takes the 8-byte-unit kernels
            import torch
            t = torch.from_numpy(np.concatenate([np.zeros(1, dtype=np.uint64), v]).view(np.int64)).cuda()
            view = engine.Ciphertext.view(t.data_ptr() + 8, T, ctx, keepalive=t)
            assert key.count_satisfied(view) == want, (L, T, "misaligned view")


@pytest.mark.parametrize("L,T,D", [(3, 400001, 5), (5, 250000, 64), (19, 90001, 300), (33, 60000, 16), (65, 40001, 1),
                                   (97, 20000, 2000), (193, 10001, 16), (513, 4000, 7), (999, 1501, 30000),
                                   (34, 60001, 40), (50, 40000, 3), (100, 20001, 16), (254, 8000, 500)])
def test_decrypt_window_walk_whole_grid(engine, oracle, L, T, D):
    """Odd L (and the even L that take the same kernel), enough blocks that every warp of the grid owns a run (and runs start at every phase of the block
    structure), sparse and dense keys, blocks that miss the key in exactly one word -- the first, the last, or one in
    the middle (the word a window boundary may cut)."""
    N = 64 * L - 1
    rng = np.random.default_rng(1000 + L)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    mask = key_mask(N, s)
    v = planted(rng, T, N, s, 0.6).reshape(T, L)
    hit = np.nonzero(mask)[0]
    spoil = rng.random(T) < 0.3
    which = hit[rng.integers(0, len(hit), T)]
    which[rng.random(T) < 0.2] = hit[0]
    which[rng.random(T) < 0.2] = hit[-1]
    rows = np.nonzero(spoil)[0]
    low = mask[which[rows]] & (~mask[which[rows]] + np.uint64(1))        # one key bit of that word
    v[rows, which[rows]] &= ~low
    v = np.ascontiguousarray(v.reshape(-1))
    want = oracle.count_satisfied(v, N, s)
    assert 0 < want < T
    ct = engine.Ciphertext.from_host(v, ctx)
    assert key.count_satisfied(ct) == want
    for form in (2, 3, 4, 6, 11):
        with _Env(CSGN_DEC_WINDOW=form, CSGN_DEC_WINDOW_ALL=1):
            assert key.count_satisfied(ct) == want, form
    with _Env(CSGN_DEC_WINDOW=0):
        assert key.count_satisfied(ct) == want, "without the window walk"
    with _Env(CSGN_DEC_WAVES=4):
        assert key.count_satisfied(ct) == want, "four waves"
    with _Env(CSGN_DEC_CTAS_PER_SM=1):
        assert key.count_satisfied(ct) == want, "one CTA per SM"


def test_decrypt_deferred_and_auto_lanes(engine, oracle):
    """A loop over operator* and decrypt, the way a user of the reference writes it, with the library placing the
    independent operations on its lanes; every result equals the sequential one."""
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(31)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    P = 12
    hosts = [(planted(rng, 300 + 7 * i, N, s), planted(rng, 200 - 3 * i, N, s)) for i in range(P)]
    wants = [oracle.count_satisfied(oracle.mul(a, b, L), N, s) for a, b in hosts]
    engine.set_auto_lanes(True)
    try:
        for rounds in range(3):
            cas = [engine.Ciphertext.from_host_ptr(h[0].ctypes.data, h[0].size // L, ctx) for h in hosts]
            cbs = [engine.Ciphertext.from_host_ptr(h[1].ctypes.data, h[1].size // L, ctx) for h in hosts]
            prods = [a * b for a, b in zip(cas, cbs)]
            res = [key.decrypt_deferred(p) for p in prods]
            fused = [key.mul_decrypt_deferred(a, b) for a, b in zip(cas, cbs)]
            # dependent chains across lanes: products of products, sums, permutes
            chain = (prods[0] + prods[1])
            chain += prods[2]
            rchain = key.decrypt_deferred(chain)
            sq = prods[3].clone()
            rsq = key.decrypt_deferred(sq)
            del cas, cbs
            assert [r.count() for r in res] == wants
            assert [r.count() for r in fused] == wants
            assert rchain.count() == wants[0] + wants[1] + wants[2]
            assert rsq.count() == wants[3]
            assert np.array_equal(prods[5].getValues(), oracle.mul(hosts[5][0], hosts[5][1], L))
            # overwrite a buffer that other lanes are still reading
            x = engine.Ciphertext.empty(prods[6].n_blocks, ctx)
            r0 = key.decrypt_deferred(prods[6])
            hosts_a6 = engine.Ciphertext.from_host(hosts[6][0], ctx)
            hosts_b6 = engine.Ciphertext.from_host(hosts[6][1], ctx)
            hosts_a6.mul_into(hosts_b6, x)
            r1 = key.decrypt_deferred(x)
            hosts_b6.mul_into(hosts_a6, engine.Ciphertext.empty(prods[6].n_blocks, ctx))
            assert r0.count() == wants[6] and r1.count() == wants[6]
            del prods, res, fused, chain, sq, x
        engine.sync()
    finally:
        engine.set_auto_lanes(False)


def test_one_upload_shared_by_every_item_of_a_batch(engine, oracle):
    """an operand still in flight on the copy stream and consumed by items on several lanes must be awaited
    by every lane, not only by the first consumer."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(41)
    ctx = engine.Context(N, D)
    T1, T2, P = 2000, 500, 8
    big = torch.from_numpy(random_blocks(rng, T1, N).view(np.int64)).pin_memory()
    smalls = [random_blocks(rng, T2, N) for _ in range(P)]
    cbs = [engine.Ciphertext.from_host(x, ctx) for x in smalls]
    want = [oracle.mul(big.numpy().view(np.uint64), x, L) for x in smalls]
    for rep in range(3):
        shared = engine.Ciphertext.from_host_ptr(big.data_ptr(), T1, ctx)      # async H2D, no sync
        prods = engine.mul_batch([shared] * P, cbs)
        for p in range(P):
            assert np.array_equal(prods[p].getValues(), want[p]), (rep, p)
        del shared, prods


def test_buffer_used_on_many_streams_then_freed(engine, oracle):
    """a buffer read on stream A, then on stream B, and freed on stream C is released only after both."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(43)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    dev = torch.device("cuda", torch.cuda.current_device())
    streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
    counts = torch.zeros(64, dtype=torch.int64, device=dev)
    torch.cuda.synchronize()
    wants = []
    for i in range(16):
        v = planted(rng, 20000, N, s, 0.01)
        wants.append(oracle.count_satisfied(v, N, s))
        ct = engine.Ciphertext.from_host(v, ctx)
        engine.set_stream(streams[0].cuda_stream)
        key.count_satisfied_async(ct, counts.data_ptr() + 8 * (2 * i))
        engine.set_stream(streams[1].cuda_stream)
        key.count_satisfied_async(ct, counts.data_ptr() + 8 * (2 * i + 1))
        engine.set_stream(streams[2].cuda_stream)
        del ct                                     # freed on the third stream; the next upload may reuse the storage
    engine.set_stream(0)
    torch.cuda.synchronize()
    got = counts.cpu().tolist()
    assert got[:32] == [w for w in wants for _ in (0, 1)]


def test_batched_upload_shared_storage(engine, oracle):
    """csgn_buf_upload_batch: n ragged operands (an empty one included) in one shared allocation and copies issued back to back;
    the views feed batch products on the library's lanes, are freed in any order, and the storage is recycled."""
    import ctypes
    import torch
    N, D = 1247, 3
    L = words_per_block(N)
    rng = np.random.default_rng(47)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    sizes_a, sizes_b = [700, 1, 333, 64, 0, 1200], [50, 900, 31, 64, 5, 2]
    hosts = [torch.from_numpy(planted(rng, max(t, 1), N, s)[:t * L].view(np.int64).copy()).pin_memory() for t in sizes_a + sizes_b]
    P = len(sizes_a)
    up = engine.UploadBatch([h.data_ptr() for h in hosts], sizes_a + sizes_b, ctx)
    want = [oracle.mul(hosts[p].numpy().view(np.uint64), hosts[P + p].numpy().view(np.uint64), L) if sizes_a[p] else
            np.zeros(0, dtype=np.uint64) for p in range(P)]
    dev = torch.device("cuda", torch.cuda.current_device())
    counts = torch.zeros(P, dtype=torch.int64, device=dev)
    for rep in range(4):
        ops = up.upload()
        # every view reads back what went in
        for i in (0, 3, P + 1):
            got = engine.Ciphertext(ctypes.c_void_p(ops[i]), ctx)
            assert np.array_equal(got.getValues(), hosts[i].numpy().view(np.uint64)), (rep, i)
            got._h = None                                   # the array below still owns the handle
        live = [p for p in range(P) if sizes_a[p]]
        ha = (ctypes.c_void_p * len(live))(*[ops[p] for p in live])
        hb = (ctypes.c_void_p * len(live))(*[ops[P + p] for p in live])
        ho = (ctypes.c_void_p * len(live))()
        counts.zero_()
        engine.mul_count_batch_async(key, None, None, counts.data_ptr(), arrays=(ha, hb, ho))
        if rep % 2:
            engine.free_handles(ops)                        # operands released while the kernels are still queued
        prods = [engine.Ciphertext(ctypes.c_void_p(ho[i]), ctx) for i in range(len(live))]
        for i, p in enumerate(live):
            assert np.array_equal(prods[i].getValues(), want[p]), (rep, p)
            assert int(counts[i].item()) == oracle.count_satisfied(want[p], N, s), (rep, p)
        if not rep % 2:
            for i in reversed(range(2 * P)):                # one by one, last first
                engine._lib().csgn_buf_free(ctypes.c_void_p(ops[i]))
        del prods
    engine.sync()


def test_lazy_sum_rope(engine, oracle):
    """csgn_concat_lazy (add as a zero-copy rope): decrypt, permute and a product with the sum on the
    left walk the segments; everything else flattens once; the operands may be freed at once; growing or overwriting
    a buffer that a sum refers to is refused."""
    N, D = 1247, 3
    L = words_per_block(N)
    rng = np.random.default_rng(53)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    parts = [planted(rng, t, N, s, 0.2) for t in (7000, 9001, 8000)]
    whole = np.concatenate(parts)
    small = planted(rng, 3, N, s, 0.7)
    perm = rng.permutation(N).astype(np.uint64)
    p = engine.Permutation(ctx, perm)

    def build():
        cts = [engine.Ciphertext.from_host(x, ctx) for x in parts]
        r = cts[0].add_lazy(cts[1]).add_lazy(cts[2])
        assert r.segments == 3 and r.n_blocks == 24001
        del cts                                     # the sum keeps the storage alive
        return r

    r = build()
    assert key.count_satisfied(r) == oracle.count_satisfied(whole, N, s)
    assert r.segments == 3                          # decrypt did not copy the parts together
    assert np.array_equal(r.applyPermutation(p).getValues(), oracle.permute_all(whole, N, perm))
    strict = r.applyPermutation(p, strict_ref_truncate=True)
    assert np.array_equal(strict.getValues(), oracle.permute_block(whole[:L], N, perm))
    cs = engine.Ciphertext.from_host(small, ctx)
    assert np.array_equal((r * cs).getValues(), oracle.mul(whole, small, L))          # lazy LEFT operand: per segment
    assert r.segments == 3
    bit, cnt, prod = key.mul_decrypt(cs, cs, out="alloc")
    assert cnt == oracle.count_satisfied(oracle.mul(small, small, L), N, s)
    bit, cnt = key.mul_decrypt(r, cs)                                                   # fused: flattens the left operand
    assert cnt == oracle.count_satisfied(oracle.mul(whole, small, L), N, s) and r.segments == 1
    r = build()
    assert np.array_equal((cs * r).getValues(), oracle.mul(small, whole, L))          # lazy RIGHT operand: flattened once
    assert r.segments == 1 and np.array_equal(r.getValues(), whole)
    r = build()
    assert np.array_equal(r.getValues(), whole) and r.segments == 1                    # download flattens
    r = build()
    assert r.checksum() == oracle.checksum(whole)
    # sums of sums, and the copy for small operands / long tails
    a, b = engine.Ciphertext.from_host(parts[0], ctx), engine.Ciphertext.from_host(parts[1], ctx)
    ab, ba = a.add_lazy(b), b.add_lazy(a)
    four = ab.add_lazy(ba)
    assert four.segments == 4
    assert key.count_satisfied(four) == 2 * oracle.count_satisfied(np.concatenate(parts[:2]), N, s)
    assert np.array_equal(four.clone().getValues(), np.concatenate([parts[0], parts[1], parts[1], parts[0]]))
    assert a.add_lazy(cs).segments == 1             # a small operand is copied
    assert np.array_equal(a.add_lazy(cs).getValues(), np.concatenate([parts[0], small]))
    tail = a
    for _ in range(40):
        tail = tail.add_lazy(b)
    assert tail.segments <= 32 and tail.n_blocks == 7000 + 40 * 9001
    assert key.count_satisfied(tail) == oracle.count_satisfied(parts[0], N, s) + 40 * oracle.count_satisfied(parts[1], N, s)
    del tail
    # a buffer that a sum refers to cannot be grown or overwritten in place
    with pytest.raises(engine.CsgnError):
        a += cs
    nine = engine.Ciphertext.empty(9, ctx)
    big9 = engine.Ciphertext.from_host(planted(rng, 7000, N, s), ctx)
    keep = big9.add_lazy(big9)                      # big9 is now referred to
    with pytest.raises(engine.CsgnError):
        big9.permute_into(p, big9)
    with pytest.raises(engine.CsgnError):
        engine.Ciphertext.from_host(planted(rng, 70, N, s), ctx).mul_into(engine.Ciphertext.from_host(planted(rng, 100, N, s), ctx), big9)
    del keep, nine
    del ab, ba, four
    a += cs                                         # released: growable again
    assert np.array_equal(a.getValues(), np.concatenate([parts[0], small]))


@pytest.mark.parametrize("N", [191, 4097, 2111, 32950, 1950, 12351])
def test_multiply_odd_L_as_double_blocks(engine, oracle, N):
    """Odd L with an even number of right-operand blocks: the multiply runs on 16-byte units over double blocks (every row
    of a staged as a_i || a_i).  Same words as the 8-byte-unit kernel and the oracle, ragged tiles and tiny shapes included."""
    L = words_per_block(N)
    assert L % 2 == 1
    rng = np.random.default_rng(N + 3)
    ctx = engine.Context(N, 2)
    shapes = [(1, 2), (7, 2), (2, 8), (37, 54), (300, 200), (3, 1000), (129, 66), (1000, 4)] if L <= 64 else [(1, 2), (9, 14), (40, 6), (3, 70)]
    for T1, T2 in shapes:
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        ca, cb = engine.Ciphertext.from_host(a, ctx), engine.Ciphertext.from_host(b, ctx)
        want = oracle.mul(a, b, L)
        assert np.array_equal((ca * cb).getValues(), want), (N, T1, T2)
        for knobs in (dict(CSGN_MUL_DOUBLE=0), dict(CSGN_MUL_U=1, CSGN_MUL_R=1), dict(CSGN_MUL_U=4, CSGN_MUL_R=7, CSGN_MUL_GRID=3),
                      dict(CSGN_MUL_TPB=96, CSGN_MUL_U=2)):
            with _Env(**knobs):
                assert np.array_equal((ca * cb).getValues(), want), (N, T1, T2, knobs)


@pytest.mark.parametrize("N,D", [(191, 2), (4097, 3), (2111, 1), (1950, 2), (12351, 2), (32950, 2)])
def test_fused_odd_L_as_double_blocks(engine, oracle, N, D):
    """Odd L with an even right operand: the fused multiply -> decrypt on 16-byte units over double blocks -- product
    words and count (two verdicts per double block), count-only form, against the oracle and against the 8-byte kernel."""
    L = words_per_block(N)
    assert L % 2 == 1
    rng = np.random.default_rng(N + 17)
    shapes = [(1, 2), (7, 2), (2, 8), (37, 54), (300, 200), (3, 1000), (129, 66)] if L <= 64 else [(1, 2), (9, 14), (40, 6), (3, 70)]
    seen = 0
    for T1, T2 in shapes:
        seen += check_fused(engine, oracle, N, D, T1, T2, rng, tag="double blocks")
        for knobs in (dict(CSGN_MUL_DOUBLE=0), dict(CSGN_MUL_U=1, CSGN_MUL_R=1), dict(CSGN_MUL_U=2, CSGN_MUL_R=16, CSGN_MUL_GRID=3),
                      dict(CSGN_MUL_TPB=96, CSGN_MUL_U=1, CSGN_MUL_R=32)):
            with _Env(**knobs):
                seen += check_fused(engine, oracle, N, D, T1, T2, rng, tag=str(knobs))
    assert seen > 0
