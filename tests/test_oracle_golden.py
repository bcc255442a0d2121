"""CPU: the C restatement (oracle/csgn_oracle.c) against the golden fixtures that
tests/golden/make_golden.py produced from the unmodified reference, and -- where the
reference build is present -- against the reference itself on fresh seeded inputs."""
import numpy as np
import pytest

from conftest import sha, unhex
from oracle.pyoracle import pad_mask, random_blocks, random_key, srand, words_per_block


def _case_ids(golden):
    return ["N%d_D%d_seed%d" % (c["N"], c["D"], c["seed"]) for c in golden["cases"]]


def test_golden_has_cases(golden):
    assert len(golden["cases"]) >= 6 and len(golden["raw_cases"]) >= 3


def test_geometry_matches_reference_context(oracle, golden):
    # src/Context.cpp:20-29
    for c in golden["cases"]:
        N, D = c["N"], c["D"]
        assert c["context"] == [N, D, oracle.lib.csgn_oracle_S(N, D), oracle.lib.csgn_oracle_words_per_block(N)]
        assert c["L"] == words_per_block(N)
        # size() formulas: src/Ciphertext.cpp:91-101, src/SecretKey.cpp:269-276
        assert c["ct_size_one_block"] == 32 + 16 * c["L"]
        assert c["sk_size"] == 16 + 8 * D


def test_encrypt_replays_reference_rand_order(oracle, golden):
    # src/SecretKey.cpp:35-80, :153-206 -- identical glibc rand() consumption
    for c in golden["cases"]:
        N, D, key = c["N"], c["D"], np.array(c["key"], dtype=np.uint64)
        for bits, seed, want in ((c["bits_a"], c["seed"], c["enc_a"]), (c["bits_b"], c["seed"] + 1, c["enc_b"])):
            srand(seed)
            got = np.concatenate([oracle.encrypt(b, N, D, key) for b in bits])
            assert np.array_equal(got, unhex(want))
            # pad bits of the last word are zero
            assert not np.any(got.reshape(-1, c["L"])[:, -1] & ~pad_mask(N))


def test_multiply_add_decrypt_against_golden(oracle, golden):
    for c in golden["cases"]:
        N, L, key = c["N"], c["L"], np.array(c["key"], dtype=np.uint64)
        a, b = unhex(c["enc_a"]), unhex(c["enc_b"])
        prod = oracle.mul(a, b, L)
        assert prod.size == c["mul_len"] and sha(prod) == c["mul_sha256"]
        assert np.array_equal(prod[:2 * L], unhex(c["mul_head"])) and np.array_equal(prod[-L:], unhex(c["mul_tail"]))
        if "mul" in c:
            assert np.array_equal(prod, unhex(c["mul"]))
        assert sha(oracle.canonical_bitlen(N, prod.size // L)) == c["mul_bitlen_sha256"]
        summ = oracle.concat(a, b)
        assert summ.size == c["add_len"] and sha(summ) == c["add_sha256"]
        assert sha(oracle.canonical_bitlen(N, summ.size // L)) == c["add_bitlen_sha256"]
        for words, want in ((a, c["dec_a"]), (b, c["dec_b"]), (prod, c["dec_mul"]), (summ, c["dec_add"])):
            assert oracle.decrypt(words, N, key) == want
            assert oracle.decrypt(words, N, key, unpacked=True) == want
        # the scheme's invariants hold on the golden bits themselves
        assert c["dec_a"] == sum(c["bits_a"]) % 2 and c["dec_b"] == sum(c["bits_b"]) % 2
        assert c["dec_mul"] == (c["dec_a"] & c["dec_b"]) and c["dec_add"] == (c["dec_a"] ^ c["dec_b"])


def test_permutation_against_golden(oracle, golden):
    for c in golden["cases"]:
        N, D, L, key = c["N"], c["D"], c["L"], np.array(c["key"], dtype=np.uint64)
        srand(c["seed"] + 2)
        perm = oracle.perm_generate(N)                       # src/Permutation.cpp:139-157
        assert sha(perm) == c["perm_sha256"] and [int(x) for x in perm[:16]] == c["perm_head"]
        inv = oracle.perm_inverse(perm)
        assert sha(inv) == c["perm_inverse_sha256"]
        assert np.array_equal(oracle.perm_compose(perm, inv), np.arange(N, dtype=np.uint64))
        assert [int(x) for x in oracle.key_permute(N, key, perm)] == c["permuted_key"]
        a = unhex(c["enc_a"])
        # reference-strict: a multi-block input yields block 0 permuted (src/Ciphertext.cpp:33-40)
        assert c["permute_strict_len"] == L
        assert np.array_equal(oracle.permute_block(a[:L], N, perm), unhex(c["permute_strict"]))
        allb = oracle.permute_all(a, N, perm)
        assert sha(allb) == c["permute_each_block_sha256"]
        if "permute_each_block" in c:
            assert np.array_equal(allb, unhex(c["permute_each_block"]))
        assert oracle.decrypt(allb, N, np.array(c["permuted_key"], dtype=np.uint64)) == c["dec_permuted"] == c["dec_a"]


def test_raw_block_cases_against_golden(oracle, golden):
    for c in golden["raw_cases"]:
        N, L = c["N"], c["L"]
        rng = np.random.default_rng(c["seed"])
        a = rng.integers(0, 2**64, size=(c["T1"], L), dtype=np.uint64)
        b = rng.integers(0, 2**64, size=(c["T2"], L), dtype=np.uint64)
        a[:, L - 1] &= pad_mask(N)
        b[:, L - 1] &= pad_mask(N)
        a, b = a.reshape(-1), b.reshape(-1)
        key = rng.permutation(N)[:c["D"]].astype(np.uint64)
        assert [int(x) for x in key] == c["key"] and sha(a) == c["a_sha256"] and sha(b) == c["b_sha256"]
        prod = oracle.mul(a, b, L)
        assert sha(prod) == c["mul_sha256"]
        assert oracle.decrypt(prod, N, key) == c["dec_mul"] and oracle.decrypt(a, N, key) == c["dec_a"]
        assert oracle.mul_checksum(a, b, L) == oracle.checksum(prod)


def test_philox_known_answers_and_batch_encrypt(oracle):
    """The counter-based generator behind csgn_encrypt_batch: Random123's published philox4x32-10 vectors,
    and the construction itself (every block decrypts to its bit; splitting a batch changes nothing)."""
    import ctypes
    u32x4, u32x2 = ctypes.c_uint32 * 4, ctypes.c_uint32 * 2
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
                           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
                           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
                            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))):
        out = u32x4()
        oracle.lib.csgn_oracle_philox4x32_10(u32x4(*ctr), u32x2(*key), out)
        assert tuple(out) == want
    rng = np.random.default_rng(4)
    for N, D in ((1247, 16), (191, 5), (63, 4), (65, 1)):
        L = words_per_block(N)
        s = rng.permutation(N)[:D].astype(np.uint64)
        bits = rng.integers(0, 2, size=200).astype(np.uint8)
        enc = oracle.encrypt_batch(bits, N, s, seed=77)
        assert [oracle.decrypt(enc[i * L:(i + 1) * L], N, s) for i in range(200)] == [int(b) for b in bits]
        assert oracle.decrypt(enc, N, s) == int(bits.sum() & 1)
        assert np.array_equal(oracle.encrypt_batch(bits[50:], N, s, seed=77, first_block=50), enc[50 * L:])
        assert not np.any(enc.reshape(200, L)[:, -1] & ~pad_mask(N))


def test_chunk_identities(oracle):
    # (A1||A2)*B = (A1*B)||(A2*B);  Dec(C1||C2) = Dec(C1)^Dec(C2)
    rng = np.random.default_rng(5)
    N, D = 1247, 3
    L = words_per_block(N)
    a, b, key = random_blocks(rng, 9, N), random_blocks(rng, 7, N), random_key(rng, N, D)
    whole = oracle.mul(a, b, L)
    parts = np.concatenate([oracle.mul(a[:4 * L], b, L), oracle.mul(a[4 * L:], b, L)])
    assert np.array_equal(whole, parts)
    assert oracle.decrypt(whole, N, key) == oracle.decrypt(parts[:10 * L], N, key) ^ oracle.decrypt(parts[10 * L:], N, key)
    assert oracle.decrypt(whole, N, key) == oracle.decrypt(a, N, key) & oracle.decrypt(b, N, key)
    assert oracle.count_satisfied(whole, N, key) == oracle.count_satisfied(a, N, key) * oracle.count_satisfied(b, N, key)


@pytest.mark.parametrize("N,D", [(1247, 16), (16383, 64), (65, 2), (191, 5), (63, 4), (1, 1), (64 * 3 + 1, 7)])
def test_oracle_matches_live_reference(oracle, ref, N, D):
    """Differential run on fresh seeds; skipped where the reference build did not travel."""
    rng = np.random.default_rng(N * 31 + D)
    L = words_per_block(N)
    for trial in range(3):
        T1, T2 = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        key = rng.permutation(N)[:D].astype(np.uint64)
        prod, bl = ref.mul(a, b, N, D)
        assert np.array_equal(prod, oracle.mul(a, b, L)) and np.array_equal(bl, oracle.canonical_bitlen(N, T1 * T2))
        assert np.array_equal(ref.mul_inplace(a, b, N, D)[0], prod)
        summ, bl = ref.add(a, b, N, D)
        assert np.array_equal(summ, oracle.concat(a, b)) and np.array_equal(bl, oracle.canonical_bitlen(N, T1 + T2))
        assert np.array_equal(ref.add_inplace(a, b, N, D)[0], summ)
        bits = rng.integers(0, 2, size=T1)
        enc = ref.encrypt_many(bits, N, D, key, seed=100 + trial)
        srand(100 + trial)
        assert np.array_equal(enc, np.concatenate([oracle.encrypt(int(x), N, D, key) for x in bits]))
        for words in (enc, prod, summ):
            assert ref.decrypt(words, N, D, key) == oracle.decrypt(words, N, key)
        if N > 1:
            perm = ref.perm_generate(N, seed=200 + trial)
            srand(200 + trial)
            assert np.array_equal(perm, oracle.perm_generate(N))
            assert np.array_equal(ref.perm_inverse(perm), oracle.perm_inverse(perm))
            assert np.array_equal(ref.key_permute(N, D, key, perm), oracle.key_permute(N, key, perm))
            strict, sbl = ref.permute(enc, N, D, perm)
            assert np.array_equal(strict, oracle.permute_block(enc[:L], N, perm))
            assert np.array_equal(sbl, oracle.canonical_bitlen(N, 1))


def test_oracle_matches_live_reference_on_random_parameters(oracle, ref):
    """The same differential run over random (N, D) -- every block geometry the fixed list above does not name
    (N % 64 == 0 is left out: the reference writes past its arrays there, SURVEY hazard 3; D <= N/3 because the
    reference's key generator scans uninitialised slots -- hazard 6 -- and can spin forever when nearly every position
    has to be drawn)."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
    @given(st.integers(6, 700).filter(lambda n: n % 64 != 0), st.integers(1, 20), st.integers(0, 2**31 - 1))
    def run(N, D, seed):
        D = max(1, min(D, N // 3))
        rng = np.random.default_rng(seed)
        L = words_per_block(N)
        T1, T2 = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        a, b = random_blocks(rng, T1, N), random_blocks(rng, T2, N)
        key = rng.permutation(N)[:D].astype(np.uint64)
        prod, bl = ref.mul(a, b, N, D)
        assert np.array_equal(prod, oracle.mul(a, b, L)) and np.array_equal(bl, oracle.canonical_bitlen(N, T1 * T2))
        summ, _ = ref.add(a, b, N, D)
        assert np.array_equal(summ, oracle.concat(a, b))
        bits = rng.integers(0, 2, size=T1)
        enc = ref.encrypt_many(bits, N, D, key, seed=seed % 1000 + 1)
        srand(seed % 1000 + 1)
        assert np.array_equal(enc, np.concatenate([oracle.encrypt(int(x), N, D, key) for x in bits]))
        # planted blocks so that the fold meets satisfied blocks too, not only the (rare) random hit
        planted = prod.copy().reshape(T1 * T2, L)
        planted[:: 2] |= oracle.key_mask(N, key)
        for words in (enc, prod, summ, planted.reshape(-1)):
            assert ref.decrypt(words, N, D, key) == oracle.decrypt(words, N, key)
            assert oracle.decrypt(words, N, key) == oracle.count_satisfied(words, N, key) & 1
        perm = ref.perm_generate(N, seed=seed % 977 + 1)
        srand(seed % 977 + 1)
        assert np.array_equal(perm, oracle.perm_generate(N))
        assert np.array_equal(ref.perm_inverse(perm), oracle.perm_inverse(perm))
        assert np.array_equal(ref.key_permute(N, D, key, perm), oracle.key_permute(N, key, perm))
        one = enc[:L]
        strict, _ = ref.permute(one, N, D, perm)
        assert np.array_equal(strict, oracle.permute_block(one, N, perm))
        # Dec_{pi(k)}(pi(c)) = Dec_k(c), block by block, in the reference and in the oracle
        pk = oracle.key_permute(N, key, perm)
        assert ref.decrypt(strict, N, D, pk) == oracle.decrypt(one, N, key)

    run()


def test_oracle_c_code_clean_under_sanitizers(tmp_path):
    """oracle/csgn_oracle.c rebuilt with -fsanitize=address,undefined and driven through every entry point on eight
    geometries (odd L, L = 1, N % 64 == 0 included): the checker itself must not rely on undefined behaviour."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    so = str(tmp_path / "libcsgn_oracle_asan.so")
    c = subprocess.run(["gcc", "-O1", "-g", "-fPIC", "-shared", "-fsanitize=address,undefined", "-fno-omit-frame-pointer",
                        "-o", so, os.path.join(root, "oracle", "csgn_oracle.c")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    assert c.returncode == 0, c.stdout[-2000:]
    libasan = subprocess.run(["gcc", "-print-file-name=libasan.so"], stdout=subprocess.PIPE, text=True).stdout.strip()
    if not os.path.isabs(libasan) or not os.path.exists(libasan):
        pytest.skip("libasan.so not found next to gcc")
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from oracle.pyoracle import Oracle, random_blocks, random_key, words_per_block, srand
o = Oracle(%r)
rng = np.random.default_rng(1)
for N, D in ((1247, 16), (16383, 64), (65, 2), (191, 5), (63, 4), (1, 1), (128, 4), (2048, 8)):
    L = words_per_block(N)
    a, b = random_blocks(rng, 7, N), random_blocks(rng, 5, N)
    s = random_key(rng, N, D)
    p = o.mul(a, b, L); o.concat(a, b); o.decrypt(p, N, s); o.count_satisfied(p, N, s); o.checksum(p); o.mul_checksum(a, b, L)
    if N > 1:
        srand(3); perm = o.perm_generate(N); o.permute_all(p, N, perm); o.permute_block(a[:L], N, perm)
        o.key_permute(N, s, perm); o.perm_inverse(perm); o.perm_compose(perm, perm)
    srand(5); e = o.encrypt(1, N, D, s); o.encrypt(0, N, D, s); o.keygen(N, D)
    o.encrypt_batch(np.array([0, 1, 1, 0], dtype=np.uint8), N, s, 99); o.bits_text(e, N)
print("oracle under ASan/UBSan: clean")
''' % (root, so)
    env = dict(os.environ, LD_PRELOAD=libasan, ASAN_OPTIONS="detect_leaks=0:protect_shadow_gap=0:halt_on_error=1",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=600)
    assert r.returncode == 0 and "clean" in r.stdout, r.stdout[-3000:]
    assert "AddressSanitizer" not in r.stdout and "runtime error" not in r.stdout, r.stdout[-3000:]
