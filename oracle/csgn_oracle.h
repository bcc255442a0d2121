/*
 * csgn_oracle.h -- CPU restatement of the certFHE/CSGN ciphertext hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load it, and there only as the checker.  The product path
 * (csgn_b200/csrc, include/csgn.h) never links or calls any of this.
 *
 * Parity status: PINNED.  The reference ships no golden vectors,
 * so the restatement is pinned against the unmodified reference itself, compiled
 * from /root/reference/src into oracle/_ref/libcertfhe_ref.so (oracle/Makefile),
 * and against tests/golden/ *.json fixtures generated from that build by
 * tests/golden/make_golden.py.
 *
 * All words are uint64_t, bits MSB-first: position p of a block lives in word
 * p>>6, bit 63-(p&63)  (reference src/SecretKey.cpp:176-197, :116-121).
 */
#ifndef CSGN_ORACLE_H_
#define CSGN_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/Context.cpp:20-29 -- words per block, valid bits of the last word (0 => 64). */
uint64_t csgn_oracle_words_per_block(uint64_t N);
uint64_t csgn_oracle_S(uint64_t N, uint64_t D);
/* src/SecretKey.cpp:171-173 -- canonical bitlen [64,...,64,rem] repeated n_blocks times. */
void csgn_oracle_canonical_bitlen(uint64_t N, uint64_t n_blocks, uint64_t *out);

/* src/Ciphertext.cpp:133-179 (and :124-131) -- all-pairs AND, i-major / j-minor. */
void csgn_oracle_mul(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2,
                     uint64_t L, uint64_t *out);
/* Same product, never stored: XOR and wrapping sum of all output words, plus the
 * order-sensitive fold sum(w[i]*(2i+1)) mod 2^64 -- for sizes that do not fit. */
void csgn_oracle_mul_checksum(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2,
                              uint64_t L, uint64_t *xor_out, uint64_t *sum_out, uint64_t *wsum_out);
/* Checksum of a stored word array with the same three folds. */
void csgn_oracle_checksum(const uint64_t *v, uint64_t n_words,
                          uint64_t *xor_out, uint64_t *sum_out, uint64_t *wsum_out);

/* src/Ciphertext.cpp:107-122 -- concatenation. */
void csgn_oracle_concat(const uint64_t *a, uint64_t n_words_a, const uint64_t *b,
                        uint64_t n_words_b, uint64_t *out);

/* Appendix A.2 -- key positions -> per-word mask. mask must hold L words. */
void csgn_oracle_key_mask(uint64_t N, const uint64_t *s, uint64_t D, uint64_t *mask);
/* src/SecretKey.cpp:104-147 -- XOR over blocks of AND over the D secret bits. */
uint64_t csgn_oracle_decrypt(const uint64_t *v, uint64_t T, uint64_t N,
                             const uint64_t *s, uint64_t D);
/* Literal variant: unpack every valid bit to one byte first, as the reference does
 * (src/SecretKey.cpp:113-124), then index values[n*k+s[i]].  Small sizes only. */
uint64_t csgn_oracle_decrypt_unpacked(const uint64_t *v, uint64_t T, uint64_t N,
                                      const uint64_t *s, uint64_t D);
/* Number of blocks whose D secret bits are all one (decrypt = count & 1). */
uint64_t csgn_oracle_count_satisfied(const uint64_t *v, uint64_t T, uint64_t N,
                                     const uint64_t *s, uint64_t D);

/* src/Ciphertext.cpp:24-69 -- out_bit[i] = in_bit[perm[i]] for one block, pad bits 0. */
void csgn_oracle_permute_block(const uint64_t *in, uint64_t N, const uint64_t *perm,
                               uint64_t *out);
/* Every block permuted (the meaningful multi-block operation). */
void csgn_oracle_permute_all(const uint64_t *in, uint64_t T, uint64_t N,
                             const uint64_t *perm, uint64_t *out);
/* src/SecretKey.cpp:226-259 -- newKey = ascending { i : perm[i] in s }. Returns count. */
uint64_t csgn_oracle_key_permute(uint64_t N, const uint64_t *s, uint64_t D,
                                 const uint64_t *perm, uint64_t *out);
/* src/Permutation.cpp:8-27 and :63-78. */
void csgn_oracle_perm_inverse(const uint64_t *perm, uint64_t n, uint64_t *out);
void csgn_oracle_perm_compose(const uint64_t *p, const uint64_t *q, uint64_t n, uint64_t *out);

/* The three below consume glibc rand() in exactly the reference's call order
 * (SURVEY.md Appendix A.4); seed with srand() first. */
/* src/SecretKey.cpp:35-80 + :171-197 -- one bit -> one packed block of L words. */
void csgn_oracle_encrypt(int bit, uint64_t N, uint64_t D, const uint64_t *s, uint64_t *out);
/* src/Permutation.cpp:139-157. */
void csgn_oracle_perm_generate(uint64_t n, uint64_t *out);
/* src/SecretKey.cpp:322-335 with the slots sentinel-initialised (the reference
 * scans uninitialised memory, so upstream keys are not reproducible). */
void csgn_oracle_keygen(uint64_t N, uint64_t D, uint64_t *s);

/* Batched encryption with a counter-based generator (not in the reference,
 * which draws from glibc rand()).  Same construction as src/SecretKey.cpp:35-80 per block:
 *   Enc(1): ones at the secret positions, random bits elsewhere;
 *   Enc(0): a random secret position h ("hole"); random bits everywhere else; the bit at h is
 *           forced to 0 if every other secret position came out 1, otherwise it is random.
 * Randomness: Philox-4x32-10 keyed by `seed`; block b (global index first_block + i), 16-byte
 * unit u of the block uses counter (b_lo, b_hi, u, 0x43534731); the hole index is
 * philox(b_lo, b_hi, 0xffffffff, 0x43534731).x % D into `s` as given.  out holds n*L words. */
void csgn_oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void csgn_oracle_encrypt_batch(const uint8_t *bits, uint64_t n, uint64_t first_block, uint64_t N,
                               const uint64_t *s, uint64_t D, uint64_t seed, uint64_t *out);

/* src/Ciphertext.cpp:185-202 -- '0'/'1' text of the valid bits; out holds T*N+1 chars. */
void csgn_oracle_bits_text(const uint64_t *v, uint64_t T, uint64_t N, char *out);

#ifdef __cplusplus
}
#endif
#endif
