// ctor_probe.cpp -- where the host time of Ciphertext(V, Bitlen, len, ctx) goes (GPU box).
//
//   usage: ctor_probe [operands = 32] [rounds = 200]
//
// Context(1247,16), 1000 blocks per operand (160 KB), the operand of the bench step.  Times, per call and with the
// device drained between rounds so that nothing waits for memory: the constructor with and without the Bitlen array
// (the difference is the validation), csgn_buf_upload_copy (staging copy + upload), csgn_buf_upload from pinned memory
// (upload alone), and a plain memcpy of the same words.
#include "certFHE.h"
#include "csgn.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

using namespace certFHE;
typedef std::chrono::steady_clock clk;

static double us(clk::time_point a, clk::time_point b, long n) { return std::chrono::duration<double>(b - a).count() * 1e6 / n; }

int main(int argc, char **argv) {
    const int P = argc > 1 ? atoi(argv[1]) : 32, rounds = argc > 2 ? atoi(argv[2]) : 200;
    const uint64_t N = 1247, D = 16, T = 1000;
    Library::initializeLibrary();
    Context ctx(N, D);
    const uint64_t L = ctx.getDefaultN(), len = T * L, rem = N % 64;
    std::vector<std::vector<uint64_t>> ops(P, std::vector<uint64_t>(len));
    for (auto &v : ops)
        for (uint64_t i = 0; i < len; ++i) v[i] = ((uint64_t)rand() << 32 | (uint64_t)rand()) & (((i % L) + 1 == L) ? ~0ull << (64 - rem) : ~0ull);
    std::vector<uint64_t> bitlen(len);
    for (uint64_t i = 0; i < len; ++i) bitlen[i] = ((i % L) + 1 == L && rem) ? rem : 64;
    void *pinned = nullptr;
    if (csgn_host_alloc((size_t)P * len * 8, &pinned) != 0) return 2;
    for (int p = 0; p < P; ++p) memcpy((char *)pinned + (size_t)p * len * 8, ops[p].data(), len * 8);
    std::vector<uint64_t> scratch((size_t)P * len);

    double t_full = 0, t_nobl = 0, t_copy = 0, t_up = 0, t_memcpy = 0, t_free = 0;
    for (int r = -20; r < rounds; ++r) {
        std::vector<Ciphertext *> cs(P);
        std::vector<csgn_buf *> bs(P);
        Library::synchronize();
        auto a = clk::now();
        for (int p = 0; p < P; ++p) cs[p] = new Ciphertext(ops[p].data(), bitlen.data(), len, ctx);
        auto b = clk::now();
        Library::synchronize();
        for (int p = 0; p < P; ++p) delete cs[p];
        Library::synchronize();
        auto c = clk::now();
        for (int p = 0; p < P; ++p) cs[p] = new Ciphertext(ops[p].data(), nullptr, len, ctx);
        auto d = clk::now();
        Library::synchronize();
        for (int p = 0; p < P; ++p) delete cs[p];
        Library::synchronize();
        auto e = clk::now();
        for (int p = 0; p < P; ++p) csgn_buf_upload_copy(ops[p].data(), T, (uint32_t)L, &bs[p]);
        auto f = clk::now();
        Library::synchronize();
        auto f2 = clk::now();
        for (int p = 0; p < P; ++p) csgn_buf_free(bs[p]);
        auto f3 = clk::now();
        Library::synchronize();
        auto g = clk::now();
        for (int p = 0; p < P; ++p) csgn_buf_upload((const uint64_t *)((char *)pinned + (size_t)p * len * 8), T, (uint32_t)L, &bs[p]);
        auto h = clk::now();
        Library::synchronize();
        for (int p = 0; p < P; ++p) csgn_buf_free(bs[p]);
        Library::synchronize();
        auto i = clk::now();
        for (int p = 0; p < P; ++p) memcpy(scratch.data() + (size_t)p * len, ops[p].data(), len * 8);
        auto j = clk::now();
        if (r >= 0) {
            t_full += us(a, b, P); t_nobl += us(c, d, P); t_copy += us(e, f, P); t_up += us(g, h, P); t_memcpy += us(i, j, P);
            t_free += us(f2, f3, P);
        }
    }
    printf("per call, us: Ciphertext(V,Bitlen,len,ctx) %.2f | without Bitlen %.2f (validation %.2f) | csgn_buf_upload_copy %.2f | "
           "csgn_buf_upload from pinned memory %.2f (staging copy %.2f) | plain memcpy of 160 KB %.2f | csgn_buf_free %.2f\n",
           t_full / rounds, t_nobl / rounds, (t_full - t_nobl) / rounds, t_copy / rounds, t_up / rounds, (t_copy - t_up) / rounds,
           t_memcpy / rounds, t_free / rounds);
    csgn_host_free(pinned);
    return 0;
}
