// cpp_e2e.cpp -- the bench step through the certFHE C++ drop-in API, end to end from host arrays.
//
//   usage: cpp_e2e [pairs = 16] [steps = 20]
//
// The shape of the reference's own tests/timings.cpp:17-72 at BASELINE.json configs[1] scale, written the way a user
// of the reference writes it -- no batch calls, no streams, no extensions:
//
//     for every pair:  Ciphertext a(V, Bitlen, len, ctx), b(...);      // host arrays, deep-copied (src/Ciphertext.cpp:344-358)
//                      Ciphertext c = a * b;                            // src/Ciphertext.cpp:231-247
//                      Plaintext  p = sk.decrypt(c);                    // src/SecretKey.cpp:208-224
//     then read every p.getValue() and compare it with the host-known truth.
//
// Context(1247,16), 1000 x 1000 blocks per pair.  Two configurations are timed: the library's defaults (fused products,
// automatic lanes, deferred Plaintext) and "one kernel per operator" (setFusedProducts(false), setAutoLanes(false)) --
// the first release's behaviour.  Prints one JSON object; bench.py reports it as `e2e_cpp`.
#include "certFHE.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace certFHE;

namespace {

struct Pair {
    std::vector<uint64_t> a, b;
    unsigned char want;     // Dec(a*b) = parity of count(a)*count(b)
};

uint64_t count_satisfied(const std::vector<uint64_t> &w, uint64_t L, const std::vector<uint64_t> &mask) {
    uint64_t n = 0;
    for (uint64_t blk = 0; blk * L < w.size(); ++blk) {
        bool ok = true;
        for (uint64_t k = 0; k < L && ok; ++k) ok = (w[blk * L + k] & mask[k]) == mask[k];
        n += ok ? 1 : 0;
    }
    return n;
}

double g_host[5];   // host seconds spent in: constructors, operator*, decrypt, destructors, getValue (CPP_E2E_BREAKDOWN=1)

double run(int steps, std::vector<Pair> &pairs, const std::vector<uint64_t> &bitlen, const Context &ctx, SecretKey &sk,
           bool *ok) {
    const uint64_t len = pairs[0].a.size();
    std::vector<Plaintext> plain(pairs.size());
    const bool breakdown = getenv("CPP_E2E_BREAKDOWN") != nullptr;
    typedef std::chrono::steady_clock clk;
    for (int i = 0; i < 5; ++i) g_host[i] = 0;
    Library::synchronize();
    const auto t0 = clk::now();
    for (int s = 0; s < steps; ++s) {
        for (size_t p = 0; p < pairs.size(); ++p) {
            if (!breakdown) {
                Ciphertext a(pairs[p].a.data(), bitlen.data(), len, ctx);
                Ciphertext b(pairs[p].b.data(), bitlen.data(), len, ctx);
                Ciphertext c = a * b;
                plain[p] = sk.decrypt(c);
                continue;
            }
            const auto u0 = clk::now();
            Ciphertext *a = new Ciphertext(pairs[p].a.data(), bitlen.data(), len, ctx);
            Ciphertext *b = new Ciphertext(pairs[p].b.data(), bitlen.data(), len, ctx);
            const auto u1 = clk::now();
            Ciphertext *c = new Ciphertext(*a * *b);
            const auto u2 = clk::now();
            plain[p] = sk.decrypt(*c);
            const auto u3 = clk::now();
            delete a; delete b; delete c;
            const auto u4 = clk::now();
            g_host[0] += std::chrono::duration<double>(u1 - u0).count();
            g_host[1] += std::chrono::duration<double>(u2 - u1).count();
            g_host[2] += std::chrono::duration<double>(u3 - u2).count();
            g_host[3] += std::chrono::duration<double>(u4 - u3).count();
        }
        const auto v0 = clk::now();
        for (size_t p = 0; p < pairs.size(); ++p)
            if (plain[p].getValue() != pairs[p].want) *ok = false;
        g_host[4] += std::chrono::duration<double>(clk::now() - v0).count();
    }
    Library::synchronize();
    const double t = std::chrono::duration<double>(clk::now() - t0).count();
    if (breakdown)
        fprintf(stderr, "host us per pair: constructors %.1f, operator* %.1f, decrypt %.1f, destructors %.1f; getValue per step %.1f us; "
                        "wall per step %.1f us\n", 1e6 * g_host[0] / (steps * pairs.size()), 1e6 * g_host[1] / (steps * pairs.size()),
                1e6 * g_host[2] / (steps * pairs.size()), 1e6 * g_host[3] / (steps * pairs.size()), 1e6 * g_host[4] / steps, 1e6 * t / steps);
    return t;
}

}  // namespace

int main(int argc, char **argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 16;
    const int steps = argc > 2 ? atoi(argv[2]) : 20;
    const uint64_t N = 1247, D = 16, T = 1000;
    try {
        Library::initializeLibrary();
        Context ctx(N, D);
        const uint64_t L = ctx.getDefaultN(), len = T * L, rem = N % 64;
        std::mt19937_64 rng(12345);
        std::vector<uint64_t> pos;
        while (pos.size() < D) {
            const uint64_t t = rng() % N;
            bool seen = false;
            for (size_t i = 0; i < pos.size(); ++i) seen |= pos[i] == t;
            if (!seen) pos.push_back(t);
        }
        SecretKey sk(ctx);
        sk.setKey(pos.data(), D);
        std::vector<uint64_t> mask(L, 0);
        for (size_t i = 0; i < pos.size(); ++i) mask[pos[i] >> 6] |= 1ull << (63 - (pos[i] & 63));
        std::vector<uint64_t> bitlen(len);
        for (uint64_t i = 0; i < len; ++i) bitlen[i] = ((i % L) + 1 == L && rem) ? rem : 64;

        std::vector<Pair> pairs(P);
        for (int p = 0; p < P; ++p) {
            std::vector<uint64_t> *ops[2] = {&pairs[p].a, &pairs[p].b};
            uint64_t cnt[2];
            for (int o = 0; o < 2; ++o) {
                std::vector<uint64_t> &w = *ops[o];
                w.resize(len);
                for (uint64_t i = 0; i < len; ++i) w[i] = rng();
                if (rem)
                    for (uint64_t blk = 0; blk < T; ++blk) w[blk * L + L - 1] &= ~0ull << (64 - rem);
                const int planted = 20 + (int)(rng() % 40);      // raw blocks almost never satisfy a D=16 key
                for (int k = 0; k < planted; ++k) {
                    const uint64_t blk = rng() % T;
                    for (uint64_t j = 0; j < L; ++j) w[blk * L + j] |= mask[j];
                }
                cnt[o] = count_satisfied(w, L, mask);
            }
            pairs[p].want = (unsigned char)((cnt[0] * cnt[1]) & 1u);
        }

        bool ok = true;
        // warm-up: the memory pool grows to the working set (16 products of 160 MB in flight on two lanes), the pinned
        // staging and the kernels are loaded -- on a fresh box the first dozen steps take tens of milliseconds each
        // -- blocks of 10 steps until a block is no longer faster than the one before it (at least 3, at most 40 blocks)
        auto warm_up = [&]() {
            double prev = 1e30;
            for (int blk = 0; blk < 40; ++blk) {
                const double t = run(10, pairs, bitlen, ctx, sk, &ok);
                if (blk >= 2 && t >= 0.9 * prev) break;
                prev = t;
            }
        };
        warm_up();
        const double t_default = run(steps, pairs, bitlen, ctx, sk, &ok);
        Library::setFusedProducts(false);
        Library::setAutoLanes(false);
        warm_up();
        const double t_eager = run(steps, pairs, bitlen, ctx, sk, &ok);
        const double blocks = (double)P * T * T * steps;
        printf("{\"value\": %.6g, \"unit\": \"blocks/s\", \"ms_per_step\": %.6g, \"pairs_per_step\": %d, \"steps\": %d, "
               "\"h2d_bytes_per_step\": %llu, \"d2h_bytes_per_step\": %d, \"checked\": %s, "
               "\"one_kernel_per_operator\": {\"value\": %.6g, \"ms_per_step\": %.6g, "
               "\"note\": \"setFusedProducts(false), setAutoLanes(false): csgn_mul then csgn_decrypt_deferred on one stream\"}, "
               "\"timing\": \"std::chrono around the loop, Library::synchronize() on both sides\", "
               "\"path\": \"host uint64 arrays -> Ciphertext(V,Bitlen,len,ctx) x2 -> operator* -> SecretKey::decrypt per pair "
               "(libcertFHE.so, defaults: fused products + automatic lanes + deferred Plaintext), every Plaintext read and "
               "checked against the host-known truth each step\"}\n",
               blocks / t_default, 1e3 * t_default / steps, P, steps, (unsigned long long)(2ull * P * len * 8), P * 8,
               ok ? "true" : "false", blocks / t_eager, 1e3 * t_eager / steps);
        return ok ? 0 : 1;
    } catch (const std::exception &e) {
        fprintf(stderr, "cpp_e2e: %s\n", e.what());
        return 2;
    }
}
