// permute.cu -- K4, permutation apply: out_bit[i] = in_bit[perm[i]], i < N, per block
// (reference src/Ciphertext.cpp:24-69).  Bits are MSB-first; the pad bits of the
// last word come out zero, as the reference's repacking loop leaves them.
//
// Roofline: HBM, 16*L bytes per block (read + write), with the integer pipe as the
// co-limiter: this is the one kernel of the path whose ALU work per byte matters.
// A per-bit gather costs ~3 integer ops per bit; the SAME permutation is applied to
// every block, so the fast kernel works bit-sliced instead (~0.6 op per bit):
//
//   1. a tile of 32 blocks is read as 32-bit word columns; thread (tile, column c) holds
//      the 32x32 bit matrix "block x bit" of that column in registers and transposes it
//      (PRMT for the 16- and 8-bit stages, SHF+LOP3 for 4/2/1): word j of the result
//      carries bit j of column c for all 32 blocks -- one *slice* per bit position;
//   2. slices go to shared memory (rows padded to 33 words: conflict-free);
//   3. the permutation is now a WORD gather: output slice (c', j) = input slice
//      slice_map[c', j] (precomputed per csgn_perm; pad positions point at a zero slot);
//   4. thread (tile, column c') gathers its 32 slices, transposes back and stores the
//      32-bit word c' of the 32 output blocks (128-byte coalesced per block row).
//
// The word-gather kernel below it (one thread builds one 64-bit output word from its 64
// source bits) covers what the tile kernel cannot hold in shared memory (N > ~54,000).
#include "kernels.cuh"
#include "launch.cuh"

#include <algorithm>

namespace csgn {
namespace {

// ---------------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers: after the call x[j] bit b == (old x[b]) bit j.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void transpose32(uint32_t (&x)[32]) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {   // swap the off-diagonal 16x16 blocks: two byte permutes
        const uint32_t a = x[k], b = x[k + 16];
        x[k] = __byte_perm(a, b, 0x5410);
        x[k + 16] = __byte_perm(a, b, 0x7632);
    }
#pragma unroll
    for (int g = 0; g < 32; g += 16)
#pragma unroll
        for (int k = g; k < g + 8; ++k) {   // 8x8 blocks
            const uint32_t a = x[k], b = x[k + 8];
            x[k] = __byte_perm(a, b, 0x6240);
            x[k + 8] = __byte_perm(a, b, 0x7351);
        }
#pragma unroll
    for (int g = 0; g < 32; g += 8)
#pragma unroll
        for (int k = g; k < g + 4; ++k) {   // 4x4 blocks
            const uint32_t a = x[k], b = x[k + 4], m = 0x0f0f0f0fu;
            x[k] = (a & m) | ((b << 4) & ~m);
            x[k + 4] = ((a >> 4) & m) | (b & ~m);
        }
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; ++k) {   // 2x2 blocks
            const uint32_t a = x[k], b = x[k + 2], m = 0x33333333u;
            x[k] = (a & m) | ((b << 2) & ~m);
            x[k + 2] = ((a >> 2) & m) | (b & ~m);
        }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {       // single bits
        const uint32_t a = x[k], b = x[k + 1], m = 0x55555555u;
        x[k] = (a & m) | ((b << 1) & ~m);
        x[k + 1] = ((a >> 1) & m) | (b & ~m);
    }
}

// slice_map layout: entry j*W + c = BYTE offset, inside a tile's slice area, of the source
// slice of output (column c, bit j), or the byte offset of the zero slot.
//
// WC > 0 fixes the words per block at compile time (40 for N=1247, 512 for N=16383): every
// one of the 32+32+32 global accesses and 32+32 shared accesses of a column then uses an
// immediate offset from one base register, and a tile that lies fully inside the ciphertext
// takes a path without per-access bounds.  That is a third of the instructions of the
// runtime-W form -- this kernel is bound by the integer/issue pipes, not by HBM.
template <int WC>
__global__ void __launch_bounds__(512)
permute_sliced_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t Wrt,
                      const uint32_t *__restrict__ slice_map, uint32_t *__restrict__ out,
                      const uint32_t tiles_per_cta, const uint64_t n_groups) {
    extern __shared__ uint32_t S[];                  // tiles_per_cta * (33*W + 1) words
    const uint32_t W = WC ? (uint32_t)WC : Wrt;
    const uint32_t tile_words = 33u * W + 1u;        // last word = the zero slot
    const uint32_t items = tiles_per_cta * W;
    pdl_enter();

    for (uint32_t t = threadIdx.x; t < tiles_per_cta; t += blockDim.x) S[t * tile_words + 33u * W] = 0u;

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint64_t tile0 = grp * tiles_per_cta;
        // ---- 1+2: columns in, transposed, slices to shared memory -----------------------
        for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
            const uint32_t tl = it / W, c = it - tl * W;
            const uint64_t blk0 = (tile0 + tl) * 32u;
            uint32_t x[32];
            const uint32_t *src = in + blk0 * W + c;
            if (blk0 + 32u <= T) {
#pragma unroll
                for (int b = 0; b < 32; ++b) x[b] = __ldcs(src + (uint32_t)b * W);
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b) x[b] = (blk0 + b < T) ? __ldcs(src + (uint64_t)b * W) : 0u;
            }
            transpose32(x);
            uint32_t *dst = S + tl * tile_words + 33u * c;
#pragma unroll
            for (int j = 0; j < 32; ++j) dst[j] = x[j];
        }
        __syncthreads();
        // ---- 3+4: gather slices, transpose back, columns out ----------------------------
        for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
            const uint32_t tl = it / W, c = it - tl * W;
            const uint64_t blk0 = (tile0 + tl) * 32u;
            const unsigned char *Sl = reinterpret_cast<const unsigned char *>(S + tl * tile_words);
            const uint32_t *map = slice_map + c;
            uint32_t y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j)
                y[j] = *reinterpret_cast<const uint32_t *>(Sl + __ldg(map + (uint32_t)j * W));
            transpose32(y);
            uint32_t *dst = out + blk0 * W + c;
            if (blk0 + 32u <= T) {
#pragma unroll
                for (int b = 0; b < 32; ++b) __stcs(dst + (uint32_t)b * W, y[b]);
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b)
                    if (blk0 + b < T) __stcs(dst + (uint64_t)b * W, y[b]);
            }
        }
        __syncthreads();   // before the next group overwrites the slices
    }
}

// Word-gather kernel: one thread builds one 64-bit output word from its 64 source
// bits.  Works for any N and L; the input block and the map are read through L1.
__global__ void __launch_bounds__(256)
permute_gather_kernel(const uint64_t *__restrict__ in, const uint64_t total_words, const uint32_t L,
                      const uint32_t N, const uint32_t *__restrict__ src_map, uint64_t *__restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    pdl_enter();
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

namespace {

template <int WC>
cudaError_t launch_sliced(const uint64_t *in, uint64_t T, uint32_t W, const uint32_t *slice_map, uint64_t *out,
                          uint32_t tiles_per_cta, uint32_t tpb, size_t smem, cudaStream_t stream) {
    const DeviceProps &dp = device_props();
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(permute_sliced_kernel<WC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)dp.smem_optin);
        if (e != cudaSuccess) return e;
        configured = dp.smem_optin;
    }
    static int per_sm_cache = 0;
    static uint32_t cache_tpb = 0;
    static size_t cache_smem = 0;
    if (per_sm_cache == 0 || cache_tpb != tpb || cache_smem != smem) {
        int per_sm = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, permute_sliced_kernel<WC>, (int)tpb, smem) !=
                cudaSuccess || per_sm < 1)
            per_sm = 1;
        per_sm_cache = per_sm;
        cache_tpb = tpb;
        cache_smem = smem;
    }
    const uint64_t n_tiles = (T + 31) / 32;
    const uint64_t n_groups = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>(n_groups, (uint64_t)dp.sm_count * per_sm_cache * (uint64_t)env_long("CSGN_PERM_WAVES", 16)));
    return launch_kernel(permute_sliced_kernel<WC>, grid, tpb, smem, stream, reinterpret_cast<const uint32_t *>(in), T, W,
                         slice_map, reinterpret_cast<uint32_t *>(out), tiles_per_cta, n_groups);
}

}  // namespace

bool permute_sliced_supported(uint32_t L) {
    const size_t need = ((size_t)33 * 2 * L + 1) * sizeof(uint32_t);
    return need <= device_props().smem_optin && 2 * L >= 1;
}

cudaError_t launch_permute(const uint64_t *in, uint64_t T, uint32_t L, uint32_t N, const uint32_t *src_map,
                           const uint32_t *slice_map, uint64_t *out, cudaStream_t stream) {
    if (T == 0 || L == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3u) == 0;
    if (slice_map && aligned4 && permute_sliced_supported(L) && !env_long("CSGN_PERM_GATHER", 0)) {
        const uint32_t W = 2 * L;
        const uint32_t tile_words = 33u * W + 1u;
        // several tiles per CTA when a block is short (N=1247: W=40 -> 4 tiles, 160 work items)
        uint32_t tiles_per_cta = std::max<uint32_t>(1, (uint32_t)env_long("CSGN_PERM_ITEMS", 160) / W);
        const uint64_t n_tiles = (T + 31) / 32;
        tiles_per_cta = (uint32_t)std::min<uint64_t>(tiles_per_cta, n_tiles);
        while (tiles_per_cta > 1 && (size_t)tiles_per_cta * tile_words * 4 > 48 * 1024) --tiles_per_cta;
        const size_t smem = (size_t)tiles_per_cta * tile_words * sizeof(uint32_t);
        const uint32_t items = tiles_per_cta * W;
        const uint32_t tpb = std::min<uint32_t>(512, (items + 31) / 32 * 32);
        count_launch();
        if (W == 40) return launch_sliced<40>(in, T, W, slice_map, out, tiles_per_cta, tpb, smem, stream);
        if (W == 512) return launch_sliced<512>(in, T, W, slice_map, out, tiles_per_cta, tpb, smem, stream);
        return launch_sliced<0>(in, T, W, slice_map, out, tiles_per_cta, tpb, smem, stream);
    }
    const uint64_t total = T * L;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((total + 255) / 256, (uint64_t)dp.sm_count * 8));
    count_launch();
    return launch_kernel(permute_gather_kernel, grid, 256, 0, stream, in, total, L, N, src_map, out);
}

}  // namespace csgn
