// permute.cu -- K4, permutation apply: out_bit[i] = in_bit[perm[i]], i < N, per block
// (reference src/Ciphertext.cpp:24-69).  Bits are MSB-first; the pad bits of the
// last word come out zero, as the reference's repacking loop leaves them.
//
// Roofline: HBM, 16*L bytes per block (read + write), with the integer pipe as the
// co-limiter: this is the one kernel of the path whose ALU work per byte matters.
//
// The permutation is the same for every block, so it is precomputed once per
// csgn_perm as a source map: src_map[i] = (perm[i]>>6)<<6 | (63-(perm[i]&63)), i.e.
// the source word of output bit i and the right-shift that brings that bit to bit 0.
#include "kernels.cuh"

#include <algorithm>

namespace csgn {
namespace {

// Word-gather kernel: one thread builds one 64-bit output word from its 64 source
// bits.  Works for any N and L; the input block and the map are read through L1.
__global__ void __launch_bounds__(256)
permute_gather_kernel(const uint64_t *__restrict__ in, const uint64_t total_words, const uint32_t L,
                      const uint32_t N, const uint32_t *__restrict__ src_map, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_words; idx += stride) {
        const uint64_t blk = idx / L;
        const uint32_t w = (uint32_t)(idx - blk * L);
        const uint64_t *row = in + blk * L;
        const uint32_t first = w * 64u;
        const uint32_t nbits = min(64u, N - first);
        const uint32_t *map = src_map + first;
        uint64_t acc = 0;
        for (uint32_t j = 0; j < nbits; ++j) {
            const uint32_t m = __ldg(map + j);
            acc |= ((__ldg(row + (m >> 6)) >> (m & 63u)) & 1ull) << (63u - j);
        }
        out[idx] = acc;
    }
}

}  // namespace

cudaError_t launch_permute(const uint64_t *in, uint64_t T, uint32_t L, uint32_t N, const uint32_t *src_map,
                           uint64_t *out, cudaStream_t stream) {
    if (T == 0 || L == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const uint64_t total = T * L;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((total + 255) / 256, (uint64_t)dp.sm_count * 8));
    permute_gather_kernel<<<grid, 256, 0, stream>>>(in, total, L, N, src_map, out);
    count_launch();
    return cudaGetLastError();
}

}  // namespace csgn
