/*
 * csgn_oracle.c -- plain-C restatement of the certFHE/CSGN ciphertext hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see csgn_oracle.h).  Parity status: PINNED against the
 * unmodified reference build in oracle/_ref and the fixtures in tests/golden/.
 *
 * Differences from the reference that are deliberate and do not change results on
 * the inputs the reference handles without undefined behaviour:
 *   - every index is 64-bit (the reference's `int` counters overflow above
 *     2^31/N blocks: src/SecretKey.cpp:110-124, src/Ciphertext.cpp:16-31,165);
 *   - the `bitlen` side array is not materialised: it is always the periodic
 *     pattern [64]*(L-1)+[rem] (src/SecretKey.cpp:171-173) and multiply copies it
 *     through unchanged (src/Ciphertext.cpp:165-176);
 *   - N % 64 == 0 is handled (the reference writes one past the end there,
 *     src/SecretKey.cpp:173).
 */
#include "csgn_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline uint64_t bit_at(const uint64_t *blk, uint64_t p) {
    /* src/SecretKey.cpp:116-121: valid bit k of word i is (v[i] >> (63-k)) & 1 */
    return (blk[p >> 6] >> (63u - (p & 63u))) & 1u;
}

static inline void set_bit(uint64_t *blk, uint64_t p, uint64_t b) {
    /* src/SecretKey.cpp:179-183: bit s of a word is shifted to position 63-s */
    blk[p >> 6] |= (b & 1u) << (63u - (p & 63u));
}

static int contains(const uint64_t *v, uint64_t n, uint64_t x) {
    /* src/Helpers.cpp:18-26 */
    for (uint64_t i = 0; i < n; i++)
        if (v[i] == x) return 1;
    return 0;
}

uint64_t csgn_oracle_words_per_block(uint64_t N) {
    /* src/Context.cpp:24-28 */
    return N / 64 + ((N % 64) ? 1 : 0);
}

uint64_t csgn_oracle_S(uint64_t N, uint64_t D) {
    /* src/Context.cpp:22 */
    return N / (2 * D);
}

void csgn_oracle_canonical_bitlen(uint64_t N, uint64_t n_blocks, uint64_t *out) {
    uint64_t L = csgn_oracle_words_per_block(N), rem = N % 64;
    for (uint64_t b = 0; b < n_blocks; b++)
        for (uint64_t k = 0; k < L; k++)
            out[b * L + k] = (k + 1 == L && rem) ? rem : 64;
}

void csgn_oracle_mul(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2,
                     uint64_t L, uint64_t *out) {
    /* src/Ciphertext.cpp:153-163: res[k + L*i*T2 + L*j] = c1[k+L*i] & c2[k+L*j];
     * the 1x1 shortcut (:137-144, :124-131) is the same formula with T1=T2=1. */
    for (uint64_t i = 0; i < T1; i++) {
        const uint64_t *ai = a + i * L;
        for (uint64_t j = 0; j < T2; j++) {
            const uint64_t *bj = b + j * L;
            uint64_t *o = out + (i * T2 + j) * L;
            for (uint64_t k = 0; k < L; k++) o[k] = ai[k] & bj[k];
        }
    }
}

/* order-sensitive fold that still parallelises: sum of w[i]*(2i+1) mod 2^64 */
static inline uint64_t wsum_term(uint64_t idx, uint64_t w) {
    return w * (2u * idx + 1u);
}

void csgn_oracle_mul_checksum(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2,
                              uint64_t L, uint64_t *xor_out, uint64_t *sum_out, uint64_t *wsum_out) {
    uint64_t x = 0, s = 0, h = 0, idx = 0;
    for (uint64_t i = 0; i < T1; i++)
        for (uint64_t j = 0; j < T2; j++)
            for (uint64_t k = 0; k < L; k++) {
                uint64_t w = a[i * L + k] & b[j * L + k];
                x ^= w;
                s += w;
                h += wsum_term(idx++, w);
            }
    if (xor_out) *xor_out = x;
    if (sum_out) *sum_out = s;
    if (wsum_out) *wsum_out = h;
}

void csgn_oracle_checksum(const uint64_t *v, uint64_t n_words,
                          uint64_t *xor_out, uint64_t *sum_out, uint64_t *wsum_out) {
    uint64_t x = 0, s = 0, h = 0;
    for (uint64_t i = 0; i < n_words; i++) {
        x ^= v[i];
        s += v[i];
        h += wsum_term(i, v[i]);
    }
    if (xor_out) *xor_out = x;
    if (sum_out) *sum_out = s;
    if (wsum_out) *wsum_out = h;
}

void csgn_oracle_concat(const uint64_t *a, uint64_t n_words_a, const uint64_t *b,
                        uint64_t n_words_b, uint64_t *out) {
    /* src/Ciphertext.cpp:111-119 */
    if (n_words_a) memcpy(out, a, n_words_a * sizeof(uint64_t));
    if (n_words_b) memcpy(out + n_words_a, b, n_words_b * sizeof(uint64_t));
}

void csgn_oracle_key_mask(uint64_t N, const uint64_t *s, uint64_t D, uint64_t *mask) {
    uint64_t L = csgn_oracle_words_per_block(N);
    memset(mask, 0, L * sizeof(uint64_t));
    for (uint64_t i = 0; i < D; i++) set_bit(mask, s[i], 1);
}

uint64_t csgn_oracle_count_satisfied(const uint64_t *v, uint64_t T, uint64_t N,
                                     const uint64_t *s, uint64_t D) {
    uint64_t L = csgn_oracle_words_per_block(N), count = 0;
    for (uint64_t k = 0; k < T; k++) {
        const uint64_t *blk = v + k * L;
        /* src/SecretKey.cpp:133-137: dec = values[n*k+s[0]] & ... & values[n*k+s[d-1]] */
        uint64_t dec = 1;
        for (uint64_t i = 0; i < D; i++) dec &= bit_at(blk, s[i]);
        count += dec;
    }
    return count;
}

uint64_t csgn_oracle_decrypt(const uint64_t *v, uint64_t T, uint64_t N,
                             const uint64_t *s, uint64_t D) {
    /* src/SecretKey.cpp:139: _dec = (dec + _dec) % 2 over all blocks */
    return csgn_oracle_count_satisfied(v, T, N, s, D) & 1u;
}

uint64_t csgn_oracle_decrypt_unpacked(const uint64_t *v, uint64_t T, uint64_t N,
                                      const uint64_t *s, uint64_t D) {
    uint64_t L = csgn_oracle_words_per_block(N), rem = N % 64;
    uint8_t *values = (uint8_t *)malloc((T * N) != 0 ? T * N : 1);
    uint64_t idx = 0;
    /* src/SecretKey.cpp:113-124: walk every word, emit bitlen[i] bytes */
    for (uint64_t w = 0; w < T * L; w++) {
        uint64_t nbits = ((w % L) + 1 == L && rem) ? rem : 64;
        for (uint64_t k = 0; k < nbits; k++) values[idx++] = (uint8_t)((v[w] >> (63u - k)) & 1u);
    }
    uint64_t acc = 0;
    for (uint64_t k = 0; k < T; k++) {
        uint64_t dec = values[N * k + s[0]];
        for (uint64_t i = 1; i < D; i++) dec &= values[N * k + s[i]];
        acc = (dec + acc) % 2;
    }
    free(values);
    return acc;
}

void csgn_oracle_permute_block(const uint64_t *in, uint64_t N, const uint64_t *perm,
                               uint64_t *out) {
    uint64_t L = csgn_oracle_words_per_block(N);
    memset(out, 0, L * sizeof(uint64_t));
    /* src/Ciphertext.cpp:33-34 then :47-69: temp2[i] = temp[perm[i]], repacked MSB-first */
    for (uint64_t i = 0; i < N; i++) set_bit(out, i, bit_at(in, perm[i]));
}

void csgn_oracle_permute_all(const uint64_t *in, uint64_t T, uint64_t N,
                             const uint64_t *perm, uint64_t *out) {
    uint64_t L = csgn_oracle_words_per_block(N);
    for (uint64_t b = 0; b < T; b++) csgn_oracle_permute_block(in + b * L, N, perm, out + b * L);
}

uint64_t csgn_oracle_key_permute(uint64_t N, const uint64_t *s, uint64_t D,
                                 const uint64_t *perm, uint64_t *out) {
    /* src/SecretKey.cpp:231-250: indicator vector, gather through perm, list ones */
    uint8_t *ind = (uint8_t *)calloc(N ? N : 1, 1);
    for (uint64_t i = 0; i < D; i++) ind[s[i]] = 1;
    uint64_t n = 0;
    for (uint64_t i = 0; i < N; i++)
        if (ind[perm[i]]) out[n++] = i;
    free(ind);
    return n;
}

void csgn_oracle_perm_inverse(const uint64_t *perm, uint64_t n, uint64_t *out) {
    /* src/Permutation.cpp:12-22: p[i] = j where permutation[j] == i */
    for (uint64_t j = 0; j < n; j++) out[perm[j]] = j;
}

void csgn_oracle_perm_compose(const uint64_t *p, const uint64_t *q, uint64_t n, uint64_t *out) {
    /* src/Permutation.cpp:70-73: result[i] = this[permB[i]] */
    for (uint64_t i = 0; i < n; i++) out[i] = p[q[i]];
}

void csgn_oracle_encrypt(int bit, uint64_t N, uint64_t D, const uint64_t *s, uint64_t *out) {
    uint64_t L = csgn_oracle_words_per_block(N);
    uint8_t *res = (uint8_t *)calloc(N ? N : 1, 1);
    if (bit & 1) {
        /* src/SecretKey.cpp:41-48: ones at secret positions, one rand() elsewhere */
        for (uint64_t i = 0; i < N; i++) res[i] = contains(s, D, i) ? 1 : (uint8_t)(rand() % 2);
    } else {
        /* src/SecretKey.cpp:49-78 */
        uint64_t hole = s[(uint64_t)rand() % D];
        uint64_t v = 0;
        int first = 1;
        for (uint64_t i = 0; i < N; i++) {
            if (i == hole) continue;
            res[i] = (uint8_t)(rand() % 2);
            if (contains(s, D, i)) {
                if (first) { v = res[i]; first = 0; }
                v &= res[i];
            }
        }
        res[hole] = (v == 1) ? 0 : (uint8_t)(rand() % 2);
    }
    memset(out, 0, L * sizeof(uint64_t));
    for (uint64_t i = 0; i < N; i++) set_bit(out, i, res[i]);
    free(res);
}

void csgn_oracle_perm_generate(uint64_t n, uint64_t *out) {
    /* src/Permutation.cpp:144-156: fill with -1, rejection-sample each slot */
    for (uint64_t i = 0; i < n; i++) out[i] = (uint64_t)-1;
    for (uint64_t i = 0; i < n; i++) {
        uint64_t r = (uint64_t)rand() % n;
        while (contains(out, n, r)) r = (uint64_t)rand() % n;
        out[i] = r;
    }
}

void csgn_oracle_keygen(uint64_t N, uint64_t D, uint64_t *s) {
    for (uint64_t i = 0; i < D; i++) s[i] = (uint64_t)-1;
    uint64_t count = 0;
    while (count < D) {
        uint64_t t = (uint64_t)rand() % N;
        if (contains(s, D, t)) continue;
        s[count++] = t;
    }
}

void csgn_oracle_bits_text(const uint64_t *v, uint64_t T, uint64_t N, char *out) {
    uint64_t L = csgn_oracle_words_per_block(N), o = 0;
    for (uint64_t b = 0; b < T; b++)
        for (uint64_t p = 0; p < N; p++) out[o++] = (char)('0' + bit_at(v + b * L, p));
    out[o] = 0;
}

/* ---- counter-based batched encryption (restated for the GPU kernel csrc/encrypt.cu) ---------- */
void csgn_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

void csgn_oracle_encrypt_batch(const uint8_t *bits, uint64_t n, uint64_t first_block, uint64_t N,
                               const uint64_t *s, uint64_t D, uint64_t seed, uint64_t *out) {
    const uint64_t L = csgn_oracle_words_per_block(N), rem = N % 64;
    const uint64_t pad = rem ? ~0ull << (64 - rem) : ~0ull;
    const uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint64_t *mask = (uint64_t *)calloc(L ? L : 1, sizeof(uint64_t));
    csgn_oracle_key_mask(N, s, D, mask);
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t b = first_block + i;
        uint64_t *blk = out + i * L;
        for (uint64_t u = 0; 2 * u < L; u++) {
            const uint32_t ctr[4] = {(uint32_t)b, (uint32_t)(b >> 32), (uint32_t)u, 0x43534731u};
            uint32_t r[4];
            csgn_oracle_philox4x32_10(ctr, key, r);
            blk[2 * u] = (uint64_t)r[0] | ((uint64_t)r[1] << 32);
            if (2 * u + 1 < L) blk[2 * u + 1] = (uint64_t)r[2] | ((uint64_t)r[3] << 32);
        }
        blk[L - 1] &= pad;
        if (bits[i] & 1) {
            for (uint64_t w = 0; w < L; w++) blk[w] |= mask[w];
        } else {
            const uint32_t ctr[4] = {(uint32_t)b, (uint32_t)(b >> 32), 0xffffffffu, 0x43534731u};
            uint32_t r[4];
            csgn_oracle_philox4x32_10(ctr, key, r);
            const uint64_t hole = s[r[0] % D];
            const uint64_t hbit = 1ull << (63u - (hole & 63u));
            int others = 1;   /* AND over the other secret positions (vacuously 1 when D == 1) */
            for (uint64_t w = 0; w < L; w++) {
                const uint64_t m = (w == (hole >> 6)) ? (mask[w] & ~hbit) : mask[w];
                if ((blk[w] & m) != m) others = 0;
            }
            if (others) blk[hole >> 6] &= ~hbit;
        }
    }
    free(mask);
}
