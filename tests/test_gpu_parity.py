"""GPU: the CUDA path, called through the C ABI, against the oracle -- bit-exact.

Small and medium sizes compare every output word with oracle/csgn_oracle.c on the
same seeded inputs and with the golden fixtures generated from the unmodified
reference; BASELINE.json's full sizes are checked through size-independent
properties (streamed checksums, multiplicativity of the satisfied-block count,
chunk identities, permute/decrypt round trips)."""
import os

import numpy as np
import pytest

from conftest import sha, unhex
from oracle.pyoracle import pad_mask, random_blocks, random_key, srand, words_per_block

pytestmark = pytest.mark.gpu


def _ct(engine, words, N, D=16):
    return engine.Ciphertext.from_host(words, engine.Context(N, D))


class _Env:
    """Temporarily set tuning knobs (read by the launchers on every call)."""

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


# ---------------------------------------------------------------------------
# K1 multiply
# ---------------------------------------------------------------------------
MUL_SHAPES = [(1, 1), (1, 7), (7, 1), (2, 2), (37, 53), (300, 200), (1000, 3), (3, 1000), (129, 65)]


@pytest.mark.parametrize("N", [1247, 16383, 65, 191, 63, 2048, 4097, 33000, 70000])
def test_mul_matches_oracle(engine, oracle, N):
    L = words_per_block(N)
    rng = np.random.default_rng(N)
    shapes = MUL_SHAPES if L <= 64 else [(1, 1), (1, 5), (5, 1), (9, 14), (40, 33)]
    for T1, T2 in shapes:
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        got = (_ct(engine, a, N) * _ct(engine, b, N)).getValues()
        assert got.size == T1 * T2 * L
        assert np.array_equal(got, oracle.mul(a, b, L)), (N, T1, T2)


@pytest.mark.parametrize("knobs", [dict(CSGN_MUL_U=1), dict(CSGN_MUL_U=2, CSGN_MUL_R=1), dict(CSGN_MUL_U=8, CSGN_MUL_R=5),
                                   dict(CSGN_MUL_U=4, CSGN_MUL_R=64, CSGN_MUL_GRID=3), dict(CSGN_MUL_TPB=160),
                                   dict(CSGN_MUL_TPB=512, CSGN_MUL_U=8), dict(CSGN_MUL_GENERIC=1)])
def test_mul_every_kernel_variant(engine, oracle, knobs):
    rng = np.random.default_rng(99)
    for N in (1247, 16383):
        L = words_per_block(N)
        for T1, T2 in ((61, 97), (5, 700), (200, 1)):
            if N == 16383:
                T1, T2 = max(1, T1 // 4), max(1, T2 // 4)
            a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
            with _Env(**knobs):
                got = (_ct(engine, a, N) * _ct(engine, b, N)).getValues()
            assert np.array_equal(got, oracle.mul(a, b, L)), (knobs, N, T1, T2)


def test_mul_inplace_and_chain(engine, oracle):
    N, L = 1247, 20
    rng = np.random.default_rng(3)
    a, b, c = (random_blocks(rng, t, N) for t in (6, 5, 4))
    x = _ct(engine, a, N)
    x *= _ct(engine, b, N)
    x *= _ct(engine, c, N)
    want = oracle.mul(oracle.mul(a, b, L), c, L)
    assert x.n_blocks == 120 and np.array_equal(x.getValues(), want)
    # bitlen is the canonical pattern the reference propagates (src/Ciphertext.cpp:165-176)
    assert np.array_equal(x.getBitlen(), oracle.canonical_bitlen(N, 120))
    assert x.size() == 32 + 16 * 120 * L


def test_out_of_memory_is_an_error_not_a_crash(engine, oracle):
    """A product that cannot fit (2e5 x 2e5 blocks = 6.4 TB) is refused with a status; the engine stays usable."""
    N, L = 1247, 20
    ctx = engine.Context(N, 16)
    big = engine.Ciphertext.empty(200000, ctx)
    with pytest.raises(engine.CsgnError) as e:
        big * big
    assert e.value.code == -6
    a, b = random_blocks(np.random.default_rng(1), 5, N), random_blocks(np.random.default_rng(2), 4, N)
    assert np.array_equal((_ct(engine, a, N) * _ct(engine, b, N)).getValues(), oracle.mul(a, b, L))


def test_mul_error_behaviour(engine):
    a = _ct(engine, np.zeros(20, dtype=np.uint64), 1247)
    b = _ct(engine, np.zeros(256, dtype=np.uint64), 16383, 64)
    with pytest.raises(engine.CsgnError) as e:
        a * b
    assert e.value.code == -5
    out = engine.Ciphertext.empty(3, engine.Context(1247, 16))
    with pytest.raises(engine.CsgnError):
        a.mul_into(a, out)


# ---------------------------------------------------------------------------
# golden fixtures from the unmodified reference
# ---------------------------------------------------------------------------
def test_golden_cases_end_to_end(engine, golden):
    for c in golden["cases"]:
        N, D, L = c["N"], c["D"], c["L"]
        ctx = engine.Context(N, D)
        key = engine.SecretKey(ctx, c["key"])
        a = engine.Ciphertext.from_host(unhex(c["enc_a"]), ctx)
        b = engine.Ciphertext.from_host(unhex(c["enc_b"]), ctx)
        prod, summ = a * b, a + b
        pv, sv = prod.getValues(), summ.getValues()
        assert pv.size == c["mul_len"] and sha(pv) == c["mul_sha256"]
        assert sha(prod.getBitlen()) == c["mul_bitlen_sha256"]
        assert sv.size == c["add_len"] and sha(sv) == c["add_sha256"]
        assert sha(summ.getBitlen()) == c["add_bitlen_sha256"]
        assert key.decrypt(a) == c["dec_a"] and key.decrypt(b) == c["dec_b"]
        assert key.decrypt(prod) == c["dec_mul"] and key.decrypt(summ) == c["dec_add"]
        assert a.size() == 32 + 16 * a.getLen() and key.size() == c["sk_size"]
        # permutation: regenerate from the fixture's seed when the full list is not stored
        if "perm" in c:
            perm = np.array(c["perm"], dtype=np.uint64)
        else:
            from oracle.pyoracle import Oracle
            srand(c["seed"] + 2)
            perm = Oracle().perm_generate(N)
            assert sha(perm) == c["perm_sha256"]
        p = engine.Permutation(ctx, perm)
        strict = a.applyPermutation(p, strict_ref_truncate=True)
        assert strict.getLen() == c["permute_strict_len"]
        assert np.array_equal(strict.getValues(), unhex(c["permute_strict"]))
        allb = a.applyPermutation(p)
        assert sha(allb.getValues()) == c["permute_each_block_sha256"]
        pkey = engine.SecretKey(ctx, c["permuted_key"])
        assert pkey.decrypt(allb) == c["dec_permuted"]


def test_golden_raw_cases(engine, golden):
    for c in golden["raw_cases"]:
        N, L = c["N"], c["L"]
        rng = np.random.default_rng(c["seed"])
        a = rng.integers(0, 2**64, size=(c["T1"], L), dtype=np.uint64)
        b = rng.integers(0, 2**64, size=(c["T2"], L), dtype=np.uint64)
        a[:, L - 1] &= pad_mask(N)
        b[:, L - 1] &= pad_mask(N)
        key_pos = rng.permutation(N)[:c["D"]].astype(np.uint64)
        ctx = engine.Context(N, c["D"])
        prod = engine.Ciphertext.from_host(a, ctx) * engine.Ciphertext.from_host(b, ctx)
        assert sha(prod.getValues()) == c["mul_sha256"]
        key = engine.SecretKey(ctx, key_pos)
        assert key.decrypt(prod) == c["dec_mul"]
        assert engine.decrypt_positions(prod, ctx, key_pos) == c["dec_mul"]


# ---------------------------------------------------------------------------
# K3 decrypt
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,D", [(1247, 1), (1247, 3), (1247, 16), (16383, 2), (16383, 64), (65, 1), (191, 2),
                                 (63, 2), (2048, 3), (4097, 2), (33000, 2), (70000, 1)])
def test_decrypt_count_matches_oracle(engine, oracle, N, D):
    rng = np.random.default_rng(N * 7 + D)
    L = words_per_block(N)
    sizes = [0, 1, 2, 31, 32, 33, 255, 256, 257, 1000, 4099] if L <= 64 else [0, 1, 31, 33, 300]
    ctx = engine.Context(N, D)
    for T in sizes:
        v = random_blocks(rng, T, N)
        s = random_key(rng, N, D)
        rng.shuffle(s)
        key = engine.SecretKey(ctx, s)
        ct = engine.Ciphertext.from_host(v, ctx)
        assert key.count_satisfied(ct) == oracle.count_satisfied(v, N, s), (N, D, T)
        assert key.decrypt(ct) == oracle.decrypt(v, N, s)
    with _Env(CSGN_DEC_GENERIC=1):
        v = random_blocks(rng, 777, N)
        s = random_key(rng, N, D)
        assert engine.SecretKey(ctx, s).count_satisfied(engine.Ciphertext.from_host(v, ctx)) == oracle.count_satisfied(v, N, s)


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13])
def test_decrypt_every_kernel_variant(engine, oracle, variant):
    """N=1247 has several tuned forms of the fold (register-streamed and the bulk-copy ring); the losing ones are
    compiled only with -DCSGN_BUILD_VARIANTS (python -m csgn_b200.build --variants)."""
    if not engine.has_variants():
        pytest.skip("library built without CSGN_BUILD_VARIANTS: one kernel per shape")
    N, D = 1247, 2
    rng = np.random.default_rng(variant)
    ctx = engine.Context(N, D)
    for T in (1, 31, 32, 33, 255, 256, 257, 300, 5000, 70001):
        v, s = random_blocks(rng, T, N), random_key(rng, N, D)
        ct, key = engine.Ciphertext.from_host(v, ctx), engine.SecretKey(ctx, s)
        with _Env(CSGN_DEC_VARIANT=variant):
            assert key.count_satisfied(ct) == oracle.count_satisfied(v, N, s), (variant, T)


def test_decrypt_real_ciphertexts(engine, oracle):
    # fresh encryptions by the oracle's reference-order encrypt; D = 16 and 64
    for N, D, n in ((1247, 16, 200), (16383, 64, 12)):
        rng = np.random.default_rng(N)
        s = rng.permutation(N)[:D].astype(np.uint64)
        bits = rng.integers(0, 2, size=n)
        srand(5)
        enc = np.concatenate([oracle.encrypt(int(x), N, D, s) for x in bits])
        ctx = engine.Context(N, D)
        key = engine.SecretKey(ctx, s)
        ct = engine.Ciphertext.from_host(enc, ctx)
        assert key.count_satisfied(ct) == int(bits.sum())
        assert key.decrypt(ct) == int(bits.sum() & 1) == oracle.decrypt(enc, N, s)
        sq = ct * ct
        assert key.count_satisfied(sq) == int(bits.sum()) ** 2
        # a wrong key position flips blocks off
        s2 = s.copy()
        s2[0] = (s2[0] + 1) % N if ((s2[0] + 1) % N) not in s else s2[0]
        assert engine.SecretKey(ctx, s2).count_satisfied(ct) == oracle.count_satisfied(enc, N, s2)


def test_decrypt_product_without_materialising(engine, oracle):
    """csgn_decrypt_product: Dec(f1*...*fn) from the factors alone equals the fold of the real product."""
    N, D = 1247, 1
    rng = np.random.default_rng(21)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    hosts = [random_blocks(rng, t, N) for t in (40, 33, 17)]
    fs = [engine.Ciphertext.from_host(h, ctx) for h in hosts]
    counts = [oracle.count_satisfied(h, N, s) for h in hosts]
    bit, count = key.decrypt_product(fs)
    assert count == counts[0] * counts[1] * counts[2] and bit == (counts[0] & counts[1] & counts[2] & 1)
    prod = (fs[0] * fs[1]) * fs[2]                       # the real 22,440-block product
    assert key.count_satisfied(prod) == count and key.decrypt(prod) == bit
    assert oracle.count_satisfied(prod.getValues(), N, s) == count
    # a product nobody could store: 10 factors of 1000 blocks = 10^30 blocks; the count saturates, the bit is exact
    big = [engine.Ciphertext.from_host(random_blocks(rng, 1000, N), ctx) for _ in range(10)]
    bit, count = key.decrypt_product(big)
    cs = [key.count_satisfied(f) for f in big]
    want = 1
    for c in cs:
        want = min(want * c, 2**64 - 1)
    assert count == want and bit == int(all(c & 1 for c in cs))


def test_key_rejects_out_of_range_position(engine):
    with pytest.raises(engine.CsgnError):
        engine.SecretKey(engine.Context(1247, 2), [5, 1247])


# ---------------------------------------------------------------------------
# K2 add
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1247, 191, 63, 16383])
def test_add_and_append(engine, oracle, N):
    rng = np.random.default_rng(N + 1)
    L = words_per_block(N)
    for T1, T2 in ((1, 1), (3, 5), (1000, 1), (1, 1000), (257, 255)):
        if L > 64:
            T1, T2 = min(T1, 40), min(T2, 40)
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        ca, cb = _ct(engine, a, N), _ct(engine, b, N)
        assert np.array_equal((ca + cb).getValues(), oracle.concat(a, b))
        ca += cb
        assert np.array_equal(ca.getValues(), oracle.concat(a, b))
    # a chain of += (geometric growth) and a self-append
    acc = _ct(engine, random_blocks(rng, 1, N), N)
    want = acc.getValues()
    for i in range(12):
        piece = random_blocks(rng, 1 + i % 3, N)
        acc += _ct(engine, piece, N)
        want = oracle.concat(want, piece)
    acc += acc
    want = oracle.concat(want, want)
    assert np.array_equal(acc.getValues(), want)
    assert np.array_equal(acc.getBitlen(), oracle.canonical_bitlen(N, want.size // L))


# ---------------------------------------------------------------------------
# K4 permute
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N", [1247, 16383, 65, 191, 63, 2048, 4097, 70000])
def test_permute_matches_oracle(engine, oracle, N):
    rng = np.random.default_rng(N + 2)
    L = words_per_block(N)
    ctx = engine.Context(N, 4)
    perm = rng.permutation(N).astype(np.uint64)
    p = engine.Permutation(ctx, perm)
    for T in ([1, 2, 31, 33, 100, 1025] if L <= 64 else [1, 3, 70]):
        v = random_blocks(rng, T, N)
        ct = engine.Ciphertext.from_host(v, ctx)
        allb = ct.applyPermutation(p).getValues()
        assert np.array_equal(allb, oracle.permute_all(v, N, perm)), (N, T)
        assert not np.any(allb.reshape(T, L)[:, -1] & ~pad_mask(N))      # pad bits stay zero
        strict = ct.applyPermutation(p, strict_ref_truncate=True).getValues()
        assert np.array_equal(strict, oracle.permute_block(v[:L], N, perm))
        with _Env(CSGN_PERM_GATHER=1):      # the word-gather kernel (what very large N falls back to)
            assert np.array_equal(ct.applyPermutation(p).getValues(), allb)
        with _Env(CSGN_PERM_ITEMS=64, CSGN_PERM_WAVES=1):
            assert np.array_equal(ct.applyPermutation(p).getValues(), allb)
    # Dec_{pi(k)}(pi(c)) = Dec_k(c)
    s = random_key(rng, N, 2)
    v = random_blocks(rng, 500 if L <= 64 else 40, N)
    ct = engine.Ciphertext.from_host(v, ctx)
    k1 = engine.SecretKey(ctx, s)
    k2 = engine.SecretKey(ctx, oracle.key_permute(N, s, perm))
    assert k1.count_satisfied(ct) == k2.count_satisfied(ct.applyPermutation(p))
    # identity and inverse round trip
    ident = engine.Permutation(ctx, np.arange(N, dtype=np.uint64))
    assert np.array_equal(ct.applyPermutation(ident).getValues(), v)
    inv = engine.Permutation(ctx, oracle.perm_inverse(perm))
    assert np.array_equal(ct.applyPermutation(p).applyPermutation(inv).getValues(), v)


@pytest.mark.parametrize("N", [1247, 16383])
def test_permute_every_kernel_variant(engine, oracle, N):
    """The tuned shapes have several forms of the bit-sliced kernel (registers vs global slice map, tiles per CTA,
    the runtime-W kernel, the bulk-copy prefetch of the next group's tiles); ragged and multi-wave sizes."""
    rng = np.random.default_rng(N + 5)
    ctx = engine.Context(N, 4)
    perm = rng.permutation(N).astype(np.uint64)
    p = engine.Permutation(ctx, perm)
    for T in ([1, 31, 32, 127, 128, 129, 255, 300, 4097, 70001] if N == 1247 else [1, 33, 700]):
        v = random_blocks(rng, T, N)
        ct = engine.Ciphertext.from_host(v, ctx)
        want = oracle.permute_all(v, N, perm)
        for variant in (0, 1, 2, 3, 4, 5, 6, 7, 8):
            for waves in (1, 3):
                with _Env(CSGN_PERM_VARIANT=variant, CSGN_PERM_WAVES=waves):
                    assert np.array_equal(ct.applyPermutation(p).getValues(), want), (N, T, variant, waves)


@pytest.mark.parametrize("N", [16383, 8191, 4097, 33000, 2111, 1247])
def test_permute_plane_kernel_forms(engine, oracle, N):
    """The plane kernel (long blocks: the slices live in the tile's own 32 x W array) in every instantiation the
    launcher may pick -- bulk-copy fed and register fed, 1..4 columns per thread, compile-time and runtime W --
    on ragged and multi-wave sizes; odd L (N = 4097, 33000 is even L = 516; 2111 has L = 33) takes the cooperative
    load for its ragged last tile."""
    rng = np.random.default_rng(N + 11)
    ctx = engine.Context(N, 4)
    perm = rng.permutation(N).astype(np.uint64)
    p = engine.Permutation(ctx, perm)
    for T in ((1, 31, 32, 33, 97, 700) if N != 1247 else (1, 31, 32, 33, 97, 127, 128, 129, 255, 257, 700, 4097, 70001)):
        v = random_blocks(rng, T, N)
        ct = engine.Ciphertext.from_host(v, ctx)
        want = oracle.permute_all(v, N, perm)
        assert np.array_equal(ct.applyPermutation(p).getValues(), want), (N, T, "default")
        for form in (range(19) if N != 1247 else range(20, 26)):        # 20..25: the group form for short blocks (W = 40)
            for waves in (1, 2):
                with _Env(CSGN_PERM_PLANE=form, CSGN_PERM_WAVES=waves):
                    assert np.array_equal(ct.applyPermutation(p).getValues(), want), (N, T, form, waves)


def test_permutation_must_be_a_bijection(engine):
    ctx = engine.Context(65, 2)
    bad = np.arange(65, dtype=np.uint64)
    bad[3] = 4
    with pytest.raises(engine.CsgnError):
        engine.Permutation(ctx, bad)


# ---------------------------------------------------------------------------
# BASELINE.json configs at full size: size-independent properties
# ---------------------------------------------------------------------------
def _full_size_properties(engine, oracle, N, D_small, T1, T2, seed):
    L = words_per_block(N)
    rng = np.random.default_rng(seed)
    a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
    ctx = engine.Context(N, D_small)
    ca, cb = engine.Ciphertext.from_host(a, ctx), engine.Ciphertext.from_host(b, ctx)
    prod = ca * cb
    assert prod.n_blocks == T1 * T2
    # (1) streamed checksum of the whole product (xor, sum, index-weighted sum)
    assert prod.checksum() == oracle.mul_checksum(a, b, L)
    # (2) sampled rows, bit-exact (i-major layout: row i = blocks [i*T2, (i+1)*T2))
    for i in sorted({0, T1 - 1, T1 // 2, int(rng.integers(0, T1))}):
        assert np.array_equal(prod.download_range(i * T2, T2), oracle.mul(a[i * L:(i + 1) * L], b, L))
    # (3) chunk identity: (A1||A2)*B = (A1*B)||(A2*B)
    h = T1 // 3
    part = engine.Ciphertext.from_host(a[:h * L], ctx) * cb
    assert part.checksum() == oracle.mul_checksum(a[:h * L], b, L)
    assert np.array_equal(part.download_range(h * T2 - 1, 1), prod.download_range(h * T2 - 1, 1))
    # (4) decrypt: the satisfied-block count is multiplicative, decrypt is its parity
    s = random_key(rng, N, D_small)
    key = engine.SecretKey(ctx, s)
    na, nb = oracle.count_satisfied(a, N, s), oracle.count_satisfied(b, N, s)
    assert key.count_satisfied(ca) == na and key.count_satisfied(cb) == nb
    assert key.count_satisfied(prod) == na * nb
    assert key.decrypt(prod) == (na & 1) & (nb & 1)
    return prod, ctx, s, na * nb


def test_config2_1000x1000_full_size(engine, oracle):
    """Context(1247,16): two 1,000-block ciphertexts -> 1M output blocks, then decrypt."""
    prod, ctx, s, count = _full_size_properties(engine, oracle, 1247, 2, 1000, 1000, 1)
    # config 3: a random Permutation applied to the 1M-block product
    srand(3)
    perm = oracle.perm_generate(1247)
    p = engine.Permutation(ctx, perm)
    permuted = prod.applyPermutation(p)
    assert permuted.n_blocks == prod.n_blocks
    k2 = engine.SecretKey(ctx, oracle.key_permute(1247, s, perm))
    assert k2.count_satisfied(permuted) == count
    # EVERY word of the permuted 1M-block product against the oracle's permutation of the product (the
    # earlier check sampled 64 blocks); the oracle walks the 10^6 blocks bit by bit in a few seconds
    want = oracle.permute_all(prod.getValues(), 1247, perm)
    assert np.array_equal(permuted.getValues(), want)
    assert permuted.checksum() == oracle.checksum(want)
    del want
    assert np.array_equal(prod.applyPermutation(p, strict_ref_truncate=True).getValues(),
                          oracle.permute_block(prod.download_range(0, 1), 1247, perm))
    # with the real D=16 on raw random blocks nothing is satisfied: decrypt is 0
    key16 = engine.SecretKey(engine.Context(1247, 16), random_key(np.random.default_rng(8), 1247, 16))
    assert key16.decrypt(prod) == 0


def test_config5_large_parameters(engine, oracle):
    """Context(16383,64): L = 256 words per block."""
    _full_size_properties(engine, oracle, 16383, 2, 300, 300, 5)


def test_product_larger_than_l2_4GB(engine, oracle):
    """5000 x 5000 blocks = 25M blocks = 4 GB: the HBM-sized point of SURVEY 8d."""
    _full_size_properties(engine, oracle, 1247, 2, 5000, 5000, 7)


def test_chain_to_1e7_blocks(engine, oracle):
    """(a*b)*d with 100-, 100- and 1000-block operands: the config-4 shape, 10^7 blocks."""
    N, L = 1247, 20
    rng = np.random.default_rng(4)
    a, b, d = random_blocks(rng, 100, N), random_blocks(rng, 100, N), random_blocks(rng, 1000, N)
    ctx = engine.Context(N, 1)
    ab = oracle.mul(a, b, L)
    x = engine.Ciphertext.from_host(a, ctx)
    x *= engine.Ciphertext.from_host(b, ctx)
    x *= engine.Ciphertext.from_host(d, ctx)
    assert x.n_blocks == 10**7
    assert x.checksum() == oracle.mul_checksum(ab, d, L)
    s = random_key(rng, N, 1)
    key = engine.SecretKey(ctx, s)
    want = oracle.count_satisfied(a, N, s) * oracle.count_satisfied(b, N, s) * oracle.count_satisfied(d, N, s)
    assert key.count_satisfied(x) == want and key.decrypt(x) == want & 1


def test_chain_to_1e8_blocks_16GB(engine, oracle):
    """1000 x 1000 x 100 blocks = 10^8 blocks = 16 GB: the single-GPU point of config 4's scaling curve."""
    N, L = 1247, 20
    rng = np.random.default_rng(40)
    a, b, d = random_blocks(rng, 1000, N), random_blocks(rng, 1000, N), random_blocks(rng, 100, N)
    ctx = engine.Context(N, 1)
    s = random_key(rng, N, 1)
    key = engine.SecretKey(ctx, s)
    x = engine.Ciphertext.from_host(a, ctx) * engine.Ciphertext.from_host(b, ctx)
    y = x * engine.Ciphertext.from_host(d, ctx)
    assert y.n_blocks == 10**8
    want = oracle.count_satisfied(a, N, s) * oracle.count_satisfied(b, N, s) * oracle.count_satisfied(d, N, s)
    assert key.count_satisfied(y) == want and key.decrypt(y) == want & 1
    # rows of the 10^8-block product, bit-exact: block (i*1000+j)*100+k = a_i & b_j & d_k
    for i, j in ((0, 0), (999, 999), (123, 456)):
        ab = oracle.mul(a[i * L:(i + 1) * L], b[j * L:(j + 1) * L], L)
        assert np.array_equal(y.download_range((i * 1000 + j) * 100, 100), oracle.mul(ab, d, L))
    # chunk identity against an independent, smaller product: rows 500..509 of a
    part = (engine.Ciphertext.from_host(a[500 * L:510 * L], ctx) * engine.Ciphertext.from_host(b, ctx)) \
        * engine.Ciphertext.from_host(d, ctx)
    assert part.checksum() == oracle.mul_checksum(oracle.mul(a[500 * L:510 * L], b, L), d, L)
    assert np.array_equal(part.download_range(999999, 1), y.download_range(500 * 100000 + 999999, 1))
    del y, x, part


@pytest.mark.parametrize("N,D", [(1247, 16), (16383, 64), (191, 5), (63, 4), (65, 1), (2048, 8)])
def test_encrypt_batch_matches_oracle(engine, oracle, N, D):
    """csgn_encrypt_batch vs the C restatement (same Philox counters), and vs the scheme itself."""
    rng = np.random.default_rng(N + D)
    L = words_per_block(N)
    ctx = engine.Context(N, D)
    s = rng.permutation(N)[:D].astype(np.uint64)
    key = engine.SecretKey(ctx, s)
    n = 5000 if L <= 64 else 300
    bits = rng.integers(0, 2, size=n).astype(np.uint8)
    ct = key.encrypt_batch(bits, seed=0xC0FFEE)
    got = ct.getValues()
    assert np.array_equal(got, oracle.encrypt_batch(bits, N, s, 0xC0FFEE))
    with _Env(CSGN_ENC_LANE=1):      # the lane-per-block form (what odd L uses)
        assert np.array_equal(key.encrypt_batch(bits, seed=0xC0FFEE).getValues(), got)
    # a batch may be split anywhere (sharding): block i depends on (seed, first_block + i) only
    tail = key.encrypt_batch(bits[n // 3:], seed=0xC0FFEE, first_block=n // 3)
    assert np.array_equal(tail.getValues(), got[(n // 3) * L:])
    assert not np.array_equal(key.encrypt_batch(bits, seed=0xC0FFEF).getValues(), got)
    # it is an encryption: the sum decrypts to the XOR, every block to its own bit, pad bits are zero
    assert key.count_satisfied(ct) == int(bits.sum()) and key.decrypt(ct) == int(bits.sum() & 1)
    blocks = got.reshape(n, L)
    for i in range(0, n, max(1, n // 40)):
        assert oracle.decrypt(blocks[i], N, s) == int(bits[i])
    assert not np.any(blocks[:, -1] & ~pad_mask(N))
    # products of GPU-made ciphertexts obey the homomorphism
    other = key.encrypt_batch(bits[:37][::-1].copy(), seed=5)
    assert key.decrypt(ct * other) == (int(bits.sum() & 1) & int(bits[:37].sum() & 1))
    # non-secret positions look uniform
    m = oracle.key_mask(N, s)
    free_bits = np.unpackbits((blocks & ~m).view(np.uint8)).sum() / (n * (N - D))
    assert 0.47 < free_bits < 0.53


def test_encrypt_batch_decrypts_under_the_reference(engine, ref):
    N, D = 1247, 16
    rng = np.random.default_rng(9)
    s = rng.permutation(N)[:D].astype(np.uint64)
    bits = rng.integers(0, 2, size=64).astype(np.uint8)
    got = engine.SecretKey(engine.Context(N, D), s).encrypt_batch(bits, seed=1).getValues()
    L = words_per_block(N)
    for i in range(64):
        assert ref.decrypt(got[i * L:(i + 1) * L], N, D, s) == int(bits[i])
    assert ref.decrypt(got, N, D, s) == int(bits.sum() & 1)


def test_save_load_round_trip(engine, oracle, tmp_path):
    """csgn_buf_save / csgn_buf_load: header + raw words through pinned staging; corruption is detected."""
    rng = np.random.default_rng(77)
    for N, D, T in ((1247, 16, 1), (1247, 16, 500000), (16383, 64, 3000), (191, 5, 77)):
        ctx = engine.Context(N, D)
        v = random_blocks(rng, T, N)
        path = tmp_path / ("ct_%d_%d.csgn" % (N, T))
        engine.Ciphertext.from_host(v, ctx).save(path)
        raw = np.fromfile(path, dtype=np.uint64)
        assert raw.size == 8 + v.size and bytes(raw[:1].tobytes()) == b"CSGNCT01"
        assert [int(x) for x in raw[1:5]] == [N, D, ctx.L, T] and int(raw[5]) == oracle.checksum(v)[0]
        assert np.array_equal(raw[8:], v)                      # the file holds the reference's words verbatim
        back = engine.Ciphertext.load(path)
        assert (back.ctx.N, back.ctx.D, back.n_blocks) == (N, D, T)
        assert np.array_equal(back.getValues(), v)
    # a flipped bit, a truncated file and a foreign file are refused
    raw = np.fromfile(path, dtype=np.uint64)
    raw[20] ^= np.uint64(1)
    raw.tofile(tmp_path / "bad.csgn")
    with pytest.raises(engine.CsgnError, match="checksum"):
        engine.Ciphertext.load(tmp_path / "bad.csgn")
    np.fromfile(path, dtype=np.uint64)[:-5].tofile(tmp_path / "short.csgn")
    with pytest.raises(engine.CsgnError, match="truncated"):
        engine.Ciphertext.load(tmp_path / "short.csgn")
    (tmp_path / "other.bin").write_bytes(b"x" * 200)
    with pytest.raises(engine.CsgnError, match="not a CSGN"):
        engine.Ciphertext.load(tmp_path / "other.bin")


def test_sharded_files_and_key_permutation_files(engine, oracle, tmp_path):
    """one ciphertext file per rank (csgn_buf_save_shard), SecretKey and Permutation files; what comes
    back decrypts and permutes exactly as what went in."""
    N, D = 1247, 4
    L = words_per_block(N)
    rng = np.random.default_rng(91)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    perm = rng.permutation(N).astype(np.uint64)
    p = engine.Permutation(ctx, perm)
    v = random_blocks(rng, 10007, N)
    mask = np.zeros(L, dtype=np.uint64)
    for pos in s:
        mask[int(pos) >> 6] |= np.uint64(1 << (63 - (int(pos) & 63)))
    v.reshape(-1, L)[rng.choice(10007, 300, replace=False)] |= mask
    world = 3
    prefix = tmp_path / "big.csgn"
    for rank in range(world):                                      # every "rank" writes its own block range
        first, count = engine.shard_range(10007, rank, world)
        engine.Ciphertext.from_host(v[first * L:(first + count) * L], ctx).save_shard(prefix, rank, world, first)
    parts, total = [], 0
    for rank in range(world):
        ct, first = engine.Ciphertext.load_shard(prefix, rank, world)
        assert first == total and (ct.ctx.N, ct.ctx.D) == (N, D)
        total += ct.n_blocks
        parts.append(ct)
    assert total == 10007
    assert np.array_equal(np.concatenate([c.getValues() for c in parts]), v)
    assert sum(key.count_satisfied(c) for c in parts) == oracle.count_satisfied(v, N, s)
    with pytest.raises(engine.CsgnError, match="not written as shard"):       # another world size: another file set
        os.replace(str(prefix) + ".shard1of3", str(prefix) + ".shard1of4")
        engine.Ciphertext.load_shard(prefix, 1, 4)
    with pytest.raises(engine.CsgnError, match="one shard of a sharded"):      # a shard is not a whole ciphertext
        engine.Ciphertext.load(str(prefix) + ".shard0of3")
    # the key and the permutation through their files
    key.save(tmp_path / "k.sk")
    p.save(tmp_path / "p.pm")
    key2 = engine.SecretKey.load(tmp_path / "k.sk")
    p2 = engine.Permutation.load(tmp_path / "p.pm", ctx)
    assert (key2.ctx.N, key2.ctx.D) == (N, D) and np.array_equal(key2.s, s) and np.array_equal(p2.p, perm)
    whole = engine.Ciphertext.from_host(v, ctx)
    assert key2.count_satisfied(whole) == oracle.count_satisfied(v, N, s)
    assert np.array_equal(whole.applyPermutation(p2).getValues(), oracle.permute_all(v, N, perm))



def test_sharded_path_world_size_1_nccl(engine, oracle):
    """The N>1 code path (block-range shard, shard-local multiply chain, NCCL all-reduce of the count)
    at world size 1 on the one GPU this box has -- same calls bench.py makes under torchrun."""
    import socket
    import torch
    import torch.distributed as dist
    from csgn_b200 import sharding
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", torch.cuda.current_device()))
    try:
        N, L = 1247, 20
        rng = np.random.default_rng(31)
        a, b, d = random_blocks(rng, 64, N), random_blocks(rng, 50, N), random_blocks(rng, 9, N)
        ctx = engine.Context(N, 1)
        s = random_key(rng, N, 1)
        key = engine.SecretKey(ctx, s)
        A = sharding.ShardedCiphertext.scatter_from_host(a, ctx, engine.Ciphertext.from_host)
        assert (A.first, A.count, A.global_blocks) == (0, 64, 64)
        P = A.mul_replicated(engine.Ciphertext.from_host(b, ctx)).mul_replicated(engine.Ciphertext.from_host(d, ctx))
        full = oracle.mul(oracle.mul(a, b, L), d, L)
        assert np.array_equal(P.local.getValues(), full)
        assert P.decrypt(key) == oracle.decrypt(full, N, s)
        counts = torch.zeros(2, dtype=torch.int64, device="cuda")
        key.count_satisfied_async(P.local, counts.data_ptr())
        key.count_satisfied_async(A.local, counts.data_ptr() + 8)
        engine.sync()
        sharding.allreduce_counts(counts)
        assert counts.tolist() == [oracle.count_satisfied(full, N, s), oracle.count_satisfied(a, N, s)]
    finally:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------
# interop: caller-owned device memory and an external stream
# ---------------------------------------------------------------------------
def test_torch_views_and_stream(engine, oracle):
    import torch
    N, L = 1247, 20
    rng = np.random.default_rng(17)
    a, b = random_blocks(rng, 50, N), random_blocks(rng, 70, N)
    ctx = engine.Context(N, 2)
    dev = torch.device("cuda", torch.cuda.current_device())
    ta = torch.from_numpy(a.view(np.int64)).to(dev)
    tb = torch.from_numpy(b.view(np.int64)).to(dev)
    tout = torch.empty(50 * 70 * L, dtype=torch.int64, device=dev)
    tcount = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    engine.set_stream(stream.cuda_stream)
    try:
        va, vb, vo = (engine.Ciphertext.from_tensor(t, ctx) for t in (ta, tb, tout))
        va.mul_into(vb, vo)
        s = random_key(rng, N, 2)
        key = engine.SecretKey(ctx, s)
        key.count_satisfied_async(vo, tcount.data_ptr())
        stream.synchronize()
    finally:
        engine.set_stream(None)
    want = oracle.mul(a, b, L)
    assert np.array_equal(tout.cpu().numpy().view(np.uint64), want)
    assert int(tcount.item()) == oracle.count_satisfied(want, N, s)
    before = engine.launch_count()
    va.mul_into(vb, vo)
    engine.sync()
    assert engine.launch_count() == before + 1


# ---------------------------------------------------------------------------
# independent ciphertexts on several streams; upload storage recycling
# ---------------------------------------------------------------------------
def test_concurrent_streams_and_upload_recycling(engine, oracle):
    """Pairs enqueued round-robin on three streams (csgn_set_stream between calls, as bench.py does): folds of
    different streams run concurrently (per-launch scratch), uploads reuse the storage of freed uploads
    (completion-event cache) while earlier consumers may still be in flight.  Every count and every product
    checksum must still match the oracle."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(99)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    P, rounds = 12, 6
    sizes = [(rng.integers(50, 400), rng.integers(50, 400)) for _ in range(P)]
    host = [[(random_blocks(rng, int(t1), N), random_blocks(rng, int(t2), N)) for (t1, t2) in sizes] for _ in range(rounds)]
    want = [[oracle.count_satisfied(oracle.mul(a, b, L), N, s) for (a, b) in row] for row in host]
    pinned = [[(torch.from_numpy(a.view(np.int64)).pin_memory(), torch.from_numpy(b.view(np.int64)).pin_memory())
               for (a, b) in row] for row in host]
    main = torch.cuda.Stream()
    streams = [main, torch.cuda.Stream(), torch.cuda.Stream()]
    counts = torch.full((rounds, P), -1, dtype=torch.int64, device="cuda")
    torch.cuda.synchronize()
    try:
        for r in range(rounds):
            for p in range(P):
                st = streams[p % 3]
                engine.set_stream(st.cuda_stream)
                ta, tb = pinned[r][p]
                ha = engine.Ciphertext.from_host_ptr(ta.data_ptr(), sizes[p][0], ctx)
                hb = engine.Ciphertext.from_host_ptr(tb.data_ptr(), sizes[p][1], ctx)
                prod = ha * hb
                key.count_satisfied_async(prod, counts[r, p].data_ptr())
                del ha, hb, prod             # storage goes back while the kernels above may still be running
        torch.cuda.synchronize()
        assert counts.tolist() == want
        # the same, but every buffer is freed while ANOTHER stream is current: the release (to the pool / the upload
        # cache, from where the next round takes storage at once) has to wait for the stream that used it last
        counts.fill_(-1)
        for r in range(rounds):
            keep = []
            for p in range(P):
                engine.set_stream(streams[p % 3].cuda_stream)
                ta, tb = pinned[r][p]
                ha = engine.Ciphertext.from_host_ptr(ta.data_ptr(), sizes[p][0], ctx)
                hb = engine.Ciphertext.from_host_ptr(tb.data_ptr(), sizes[p][1], ctx)
                prod = ha * hb
                key.count_satisfied_async(prod, counts[r, p].data_ptr())
                keep.append((ha, hb, prod))
            engine.set_stream(streams[(r + 1) % 3].cuda_stream)
            del keep, ha, hb, prod
        torch.cuda.synchronize()
    finally:
        engine.set_stream(None)
    assert counts.tolist() == want
    # a recycled upload never shows stale words: upload, free, upload something else of the same size, download
    a1, a2 = random_blocks(rng, 300, N), random_blocks(rng, 300, N)
    c1 = engine.Ciphertext.from_host(a1, ctx)
    del c1
    engine.sync()
    c2 = engine.Ciphertext.from_host(a2, ctx)
    assert np.array_equal(c2.getValues(), a2)
    c2 += engine.Ciphertext.from_host(a1, ctx)      # growing an upload moves it out of the recycled storage
    assert np.array_equal(c2.getValues(), np.concatenate([a2, a1]))


def test_batch_entry_points(engine, oracle):
    """csgn_mul_batch / csgn_mul_into_batch / csgn_decrypt_count_batch_async / csgn_decrypt_batch: n independent items
    in one call, spread over the library's lanes; results as from n single calls."""
    import torch
    N, D, L = 1247, 2, 20
    rng = np.random.default_rng(4242)
    ctx = engine.Context(N, D)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    for n in (1, 2, 5, 16):
        host = [(random_blocks(rng, int(rng.integers(1, 300)), N), random_blocks(rng, int(rng.integers(1, 300)), N))
                for _ in range(n)]
        a = [engine.Ciphertext.from_host(x, ctx) for x, _ in host]
        b = [engine.Ciphertext.from_host(y, ctx) for _, y in host]
        want = [oracle.mul(x, y, L) for x, y in host]
        prods = engine.mul_batch(a, b)
        for p, w in zip(prods, want):
            assert np.array_equal(p.getValues(), w)
        outs = [engine.Ciphertext.empty(p.n_blocks, ctx) for p in prods]
        engine.mul_into_batch(a, b, outs)
        for p, w in zip(outs, want):
            assert np.array_equal(p.getValues(), w)
        counts = [oracle.count_satisfied(w, N, s) for w in want]
        dev = torch.full((n,), -1, dtype=torch.int64, device="cuda")
        key.count_satisfied_batch_async(outs, dev.data_ptr())
        engine.sync()
        assert dev.tolist() == counts
        bits, cnts = key.decrypt_batch(prods)
        assert cnts == counts and bits == [c & 1 for c in counts]
        # work enqueued after a batch on the current stream sees its results (the lanes are joined back)
        again = prods[0] * b[0] if prods[0].n_blocks * b[0].n_blocks < 200000 else None
        if again is not None:
            assert np.array_equal(again.getValues(), oracle.mul(want[0], host[0][1], L))
    # an item that fails stops the batch with an error and leaks nothing
    bad = engine.Ciphertext.from_host(random_blocks(rng, 3, 191), engine.Context(191, 2))
    with pytest.raises(engine.CsgnError, match="words per block"):
        engine.mul_batch([a[0], a[1]], [b[0], bad])
    with pytest.raises(engine.CsgnError, match="words per block"):
        key.decrypt_batch([a[0], bad])


def test_sanitize_case_sweep_runs_clean():
    """tools/sanitize_case.py -- the script meant for compute-sanitizer (closed on this pool) -- as a plain run: every
    kernel form at ragged sizes, each result against the oracle, in a fresh process."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "sanitize_case.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=900, cwd=root)
    assert r.returncode == 0 and "sanitize_case OK" in r.stdout, r.stdout[-3000:]
