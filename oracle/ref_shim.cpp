// ref_shim.cpp -- extern "C" doorway into the UNMODIFIED reference build.
//
// TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's
// own sources, taken where they lie under /root/reference/src (never copied into
// this repository), into oracle/_ref/libcertfhe_ref.so by oracle/Makefile.  The
// reference is built with -DcertFHE=certFHE_ref so that its namespace cannot
// collide with the drop-in libcertFHE of this repository inside one process.
//
// Every entry point goes through the reference's PUBLIC class API
// (Ciphertext(V,Bitlen,len,ctx), operator+ / operator*, SecretKey::decrypt,
// applyPermutation, Permutation(...)), so what is checked -- and timed -- is what
// a user of the reference gets.
#include "certFHE.h"  // -I/root/reference/src

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <random>
#include <thread>
#include <vector>

using namespace certFHE;  // expands to certFHE_ref

namespace {

std::vector<uint64_t> canonical_bitlen(uint64_t N, uint64_t n_words) {
    uint64_t L = N / 64 + (N % 64 ? 1 : 0), rem = N % 64;
    std::vector<uint64_t> bl(n_words);
    for (uint64_t i = 0; i < n_words; i++) bl[i] = ((i % L) + 1 == L && rem) ? rem : 64;
    return bl;
}

SecretKey *make_key(const Context &ctx, const uint64_t *s, uint64_t D) {
    // The constructor reseeds rand() with time (src/SecretKey.cpp:311-312) and draws
    // an unreproducible key; callers install their own and must srand() afterwards.
    SecretKey *sk = new SecretKey(ctx);
    sk->setKey(const_cast<uint64_t *>(s), D);
    return sk;
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

extern "C" {

// ---- opaque handles -------------------------------------------------------
void *ref_ct_new(const uint64_t *words, uint64_t n_words, uint64_t N, uint64_t D) {
    Context ctx(N, D);
    std::vector<uint64_t> bl = canonical_bitlen(N, n_words);
    return new Ciphertext(words, bl.data(), n_words, ctx);
}
void ref_ct_free(void *ct) { delete static_cast<Ciphertext *>(ct); }
uint64_t ref_ct_len(void *ct) { return static_cast<Ciphertext *>(ct)->getLen(); }
void ref_ct_words(void *ct, uint64_t *out) {
    Ciphertext *c = static_cast<Ciphertext *>(ct);
    std::memcpy(out, c->getValues(), c->getLen() * sizeof(uint64_t));
}
void ref_ct_bitlen(void *ct, uint64_t *out) {
    Ciphertext *c = static_cast<Ciphertext *>(ct);
    std::memcpy(out, c->getBitlen(), c->getLen() * sizeof(uint64_t));
}
long ref_ct_size(void *ct) { return static_cast<Ciphertext *>(ct)->size(); }
void *ref_ct_mul(void *a, void *b) {
    return new Ciphertext(*static_cast<Ciphertext *>(a) * *static_cast<Ciphertext *>(b));
}
void *ref_ct_add(void *a, void *b) {
    return new Ciphertext(*static_cast<Ciphertext *>(a) + *static_cast<Ciphertext *>(b));
}
void ref_ct_mul_inplace(void *a, void *b) {
    *static_cast<Ciphertext *>(a) *= *static_cast<Ciphertext *>(b);
}
void ref_ct_add_inplace(void *a, void *b) {
    *static_cast<Ciphertext *>(a) += *static_cast<Ciphertext *>(b);
}
void *ref_ct_permute(void *a, const uint64_t *perm, uint64_t n) {
    Permutation p(perm, n);
    return new Ciphertext(static_cast<Ciphertext *>(a)->applyPermutation(p));
}

void *ref_sk_new(uint64_t N, uint64_t D, const uint64_t *s) {
    Context ctx(N, D);
    return make_key(ctx, s, D);
}
void ref_sk_free(void *sk) { delete static_cast<SecretKey *>(sk); }
long ref_sk_size(void *sk) { return static_cast<SecretKey *>(sk)->size(); }
int ref_sk_decrypt(void *sk, void *ct) {
    Plaintext p = static_cast<SecretKey *>(sk)->decrypt(*static_cast<Ciphertext *>(ct));
    return p.getValue();
}
void *ref_sk_encrypt(void *sk, int bit) {
    Plaintext p(bit);
    return new Ciphertext(static_cast<SecretKey *>(sk)->encrypt(p));
}
void ref_sk_permute(void *sk, const uint64_t *perm, uint64_t n, uint64_t *out) {
    Permutation p(perm, n);
    SecretKey k2 = static_cast<SecretKey *>(sk)->applyPermutation(p);
    std::memcpy(out, k2.getKey(), k2.getLength() * sizeof(uint64_t));
}

// ---- one-shot helpers -----------------------------------------------------
void ref_perm_generate(uint64_t n, uint64_t *out) {
    Permutation p(n);  // consumes rand()
    std::memcpy(out, p.getPermutation(), n * sizeof(uint64_t));
}
void ref_perm_inverse(const uint64_t *perm, uint64_t n, uint64_t *out) {
    Permutation p(perm, n);
    Permutation q = p.getInverse();
    std::memcpy(out, q.getPermutation(), n * sizeof(uint64_t));
}
uint64_t ref_perm_compose(const uint64_t *a, uint64_t na, const uint64_t *b, uint64_t nb,
                          uint64_t *out) {
    Permutation p(a, na), q(b, nb);
    Permutation r = p + q;
    if (r.getLength()) std::memcpy(out, r.getPermutation(), r.getLength() * sizeof(uint64_t));
    return r.getLength();
}
void ref_context(uint64_t N, uint64_t D, uint64_t *out4) {
    Context c(N, D);
    out4[0] = c.getN(); out4[1] = c.getD(); out4[2] = c.getS(); out4[3] = c.getDefaultN();
}

// ---- timing of the public calls (CPU baseline for bench.py) ---------------
// `threads` independent replicas, each with its own seeded raw operands of T1 and
// T2 blocks; every replica runs `c = a*b` then `sk.decrypt(c)` once per rep.
// Harness-level parallelism only: the reference itself is single-threaded.
// Returns 0 on success; best-of-reps wall time over all replicas in *mul_s, *dec_s.
int ref_bench_mul_decrypt(uint64_t N, uint64_t D, uint64_t T1, uint64_t T2, int threads,
                          int reps, uint64_t seed, double *mul_s, double *dec_s,
                          uint64_t *parity_xor) {
    if (threads < 1 || reps < 1) return 1;
    uint64_t L = N / 64 + (N % 64 ? 1 : 0), rem = N % 64;
    uint64_t padmask = rem ? ~0ull << (64 - rem) : ~0ull;
    Context ctx(N, D);
    std::vector<uint64_t> s(D);
    {
        std::mt19937_64 g(seed ^ 0x5eedull);
        std::vector<uint64_t> pool(N);
        for (uint64_t i = 0; i < N; i++) pool[i] = i;
        for (uint64_t i = 0; i < D; i++) {
            uint64_t j = i + g() % (N - i);
            std::swap(pool[i], pool[j]);
            s[i] = pool[i];
        }
    }
    std::vector<Ciphertext *> A(threads), B(threads);
    std::vector<SecretKey *> K(threads);
    for (int t = 0; t < threads; t++) {
        std::mt19937_64 g(seed + 1000003ull * (uint64_t)t);
        std::vector<uint64_t> wa(T1 * L), wb(T2 * L);
        for (uint64_t i = 0; i < wa.size(); i++) wa[i] = g() & (((i % L) + 1 == L) ? padmask : ~0ull);
        for (uint64_t i = 0; i < wb.size(); i++) wb[i] = g() & (((i % L) + 1 == L) ? padmask : ~0ull);
        std::vector<uint64_t> bla = canonical_bitlen(N, wa.size()), blb = canonical_bitlen(N, wb.size());
        A[t] = new Ciphertext(wa.data(), bla.data(), wa.size(), ctx);
        B[t] = new Ciphertext(wb.data(), blb.data(), wb.size(), ctx);
        K[t] = make_key(ctx, s.data(), D);
    }
    double best_mul = 1e300, best_dec = 1e300;
    uint64_t parity = 0;
    for (int r = 0; r < reps; r++) {
        std::vector<Ciphertext *> C(threads, nullptr);
        std::vector<int> bits(threads, 0);
        std::atomic<int> go(0);
        auto run = [&](int phase) {
            std::vector<std::thread> th;
            go.store(0);
            for (int t = 0; t < threads; t++)
                th.emplace_back([&, t]() {
                    while (!go.load()) {}
                    if (phase == 0) C[t] = new Ciphertext(*A[t] * *B[t]);
                    else bits[t] = K[t]->decrypt(*C[t]).getValue();
                });
            double t0 = now_s();
            go.store(1);
            for (auto &x : th) x.join();
            return now_s() - t0;
        };
        double tm = run(0), td = run(1);
        if (tm < best_mul) best_mul = tm;
        if (td < best_dec) best_dec = td;
        parity = 0;
        for (int t = 0; t < threads; t++) { parity ^= (uint64_t)bits[t] << (t & 63); delete C[t]; }
    }
    for (int t = 0; t < threads; t++) { delete A[t]; delete B[t]; delete K[t]; }
    if (mul_s) *mul_s = best_mul;
    if (dec_s) *dec_s = best_dec;
    if (parity_xor) *parity_xor = parity;
    return 0;
}

}  // extern "C"
