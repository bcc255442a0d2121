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
//   2. slices go to shared memory with 128-bit stores (rows padded to kPermSliceStride = 36
//      words: 16-byte aligned and conflict-free per quarter-warp);
//   3. the permutation is now a WORD gather: output slice (c', j) = input slice
//      slice_map[c', j] (precomputed per csgn_perm; pad positions point at a zero slot);
//   4. thread (tile, column c') gathers its 32 slices, transposes back and stores the
//      32-bit word c' of the 32 output blocks (128-byte coalesced per block row).
//
// The word-gather kernel below it (one thread builds one 64-bit output word from its 64
// source bits) covers what the tile kernel cannot hold in shared memory (N > ~54,000).
#include "bulk.cuh"
#include "kernels.cuh"
#include "launch.cuh"

#include <algorithm>

namespace csgn {
namespace {

// ---------------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers: after the call x[j] bit b == (old x[b]) bit j.
//
// bitsel<M>(a, b) = (a & M) | (b & ~M) as ONE LOP3: written as `(a & m) | (b & ~m)` in C++ the
// two masks reach ptxas as two unrelated immediates (four inputs) and every output costs two
// LOP3 -- 192 of the ~620 integer-pipe instructions of a column, in a kernel bound by that pipe.
// ---------------------------------------------------------------------------------------
template <uint32_t M>
__device__ __forceinline__ uint32_t bitsel(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "n"(M));   // (a & c) | (b & ~c)
    return d;
}

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
            const uint32_t a = x[k], b = x[k + 4];
            x[k] = bitsel<0x0f0f0f0fu>(a, b << 4);
            x[k + 4] = bitsel<0x0f0f0f0fu>(a >> 4, b);
        }
#pragma unroll
    for (int g = 0; g < 32; g += 4)
#pragma unroll
        for (int k = g; k < g + 2; ++k) {   // 2x2 blocks
            const uint32_t a = x[k], b = x[k + 2];
            x[k] = bitsel<0x33333333u>(a, b << 2);
            x[k + 2] = bitsel<0x33333333u>(a >> 2, b);
        }
#pragma unroll
    for (int k = 0; k < 32; k += 2) {       // single bits
        const uint32_t a = x[k], b = x[k + 1];
        x[k] = bitsel<0x55555555u>(a, b << 1);
        x[k + 1] = bitsel<0x55555555u>(a >> 1, b);
    }
}

// slice_map layout: entry j*W + c = BYTE offset, inside a tile's slice area, of the source
// slice of output (column c, bit j), or the byte offset of the zero slot (4*kPermSliceStride*W).
constexpr uint32_t kStride = kPermSliceStride;

__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ void store_slices(uint32_t *dst, const uint32_t (&x)[32]) {
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);      // rows are 16-byte aligned (stride 36 words)
#pragma unroll
    for (int q = 0; q < 8; ++q) d4[q] = make_uint4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
}

// ---- phase A: one 32-bit column of 32 blocks in, transposed, 32 slices to shared memory ----
// FULL = the tile lies inside the ciphertext: 32 loads at immediate offsets from one base, no
// bounds.  The ragged form is a separate, never-inlined function so that its 64-bit compares do
// not leak predicates into the common path (they did: 140 ISETP per column, a fifth of the
// integer work of a kernel that is bound by the integer pipe).
template <int WC, bool FULL>
__device__ __forceinline__ void column_in(const uint32_t *__restrict__ src, uint64_t blk0, uint64_t T, uint32_t W,
                                          uint32_t *dst) {
    uint32_t x[32];
    if (FULL) {
#pragma unroll
        for (int b = 0; b < 32; ++b) x[b] = __ldcs(src + (uint32_t)b * (WC ? (uint32_t)WC : W));
    } else {
#pragma unroll
        for (int b = 0; b < 32; ++b) x[b] = (blk0 + b < T) ? __ldcs(src + (uint64_t)b * W) : 0u;
    }
    transpose32(x);
    store_slices(dst, x);
}
template <int WC>
__device__ __noinline__ void column_in_ragged(const uint32_t *__restrict__ src, uint64_t blk0, uint64_t T, uint32_t W,
                                              uint32_t *dst) {
    column_in<WC, false>(src, blk0, T, W, dst);
}

// ---- phase B: transpose the 32 gathered slices back, one column of 32 blocks out -------------
template <int WC, bool FULL>
__device__ __forceinline__ void column_out(uint32_t (&y)[32], uint32_t *__restrict__ dst, uint64_t blk0, uint64_t T,
                                           uint32_t W) {
    transpose32(y);
    if (FULL) {
#pragma unroll
        for (int b = 0; b < 32; ++b) __stcs(dst + (uint32_t)b * (WC ? (uint32_t)WC : W), y[b]);
    } else {
#pragma unroll
        for (int b = 0; b < 32; ++b)
            if (blk0 + b < T) __stcs(dst + (uint64_t)b * W, y[b]);
    }
}
// The ragged form gathers for itself (through the map in global memory): handing it the caller's
// register array by reference would force that array into local memory on the common path too.
template <int WC>
__device__ __noinline__ void column_out_ragged(uint32_t tile_addr, const uint32_t *__restrict__ map, uint32_t *__restrict__ dst,
                                               uint64_t blk0, uint64_t T, uint32_t W) {
    uint32_t y[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) y[j] = lds_u32(tile_addr + __ldg(map + (uint32_t)j * W));
    column_out<WC, false>(y, dst, blk0, T, W);
}

// Fixed-shape kernel: WC words per block and TILES tiles per CTA at compile time, WC*TILES
// threads; thread (g, c) owns column c of the CTA's tile slot g for the whole launch.  The
// CTA is persistent over groups of TILES tiles.  HOIST keeps the thread's 32 gather addresses
// (shared-window byte addresses, tile slot included) in registers for the whole launch: the
// gather is then 32 LDS and nothing else.
template <int WC, int TILES, bool HOIST, int MINB>
__global__ void __launch_bounds__(WC *TILES, MINB)
permute_fixed_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t *__restrict__ slice_map,
                     uint32_t *__restrict__ out, const uint64_t n_groups) {
    extern __shared__ __align__(128) uint32_t S[];
    constexpr uint32_t tile_words = kStride * WC + 4u;     // last 4 words = the zero slot (keeps tiles 16-byte aligned)
    const uint32_t g = threadIdx.x / WC, c = threadIdx.x - g * WC;
    uint32_t *tile = S + g * tile_words;
    const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(tile);
    // the slice map was uploaded and synchronised when the csgn_perm was created: reading it does
    // not depend on the previous kernel, so it is staged before the PDL wait
    uint32_t addr[HOIST ? 32 : 1];
    if (HOIST) {
#pragma unroll
        for (int j = 0; j < 32; ++j) addr[j] = tile_addr + __ldg(slice_map + (uint32_t)j * WC + c);
    }
    if (c == 0) tile[kStride * WC] = 0u;
    pdl_enter();

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint64_t blk0 = (grp * TILES + g) * 32u;
        const bool full = blk0 + 32u <= T;
        const uint32_t *src = in + blk0 * WC + c;
        if (full) column_in<WC, true>(src, blk0, T, WC, tile + kStride * c);
        else column_in_ragged<WC>(src, blk0, T, WC, tile + kStride * c);
        __syncthreads();
        uint32_t *dst = out + blk0 * WC + c;
        if (full) {
            uint32_t y[32];
            if (HOIST) {
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = lds_u32(addr[j]);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = lds_u32(tile_addr + __ldg(slice_map + (uint32_t)j * WC + c));
            }
            column_out<WC, true>(y, dst, blk0, T, WC);
        } else {
            column_out_ragged<WC>(tile_addr, slice_map + c, dst, blk0, T, WC);
        }
        __syncthreads();   // before the next group overwrites the slices
    }
}

// The same kernel with phase A fed from shared memory: one thread keeps a bulk asynchronous copy
// (cp.async.bulk -> the TMA engine, completion counted on an mbarrier) of the NEXT group's tiles in
// flight -- a group's TILES tiles are one contiguous run of TILES*32*W words -- while the CTA works
// on the current group.  The input of a group is then in flight across the two block-wide barriers
// and costs no registers; phase A reads its column from the raw tile (conflict-free: lanes are
// consecutive words of a row).  NBUF = 2: the copy is issued a whole iteration ahead; NBUF = 1: after
// phase A has consumed the buffer, i.e. half an iteration ahead, for less shared memory per CTA.
template <int WC, int TILES, int NBUF, int MINB>
__global__ void __launch_bounds__(WC *TILES, MINB)
permute_prefetch_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t *__restrict__ slice_map,
                        uint32_t *__restrict__ out, const uint64_t n_groups) {
    extern __shared__ __align__(128) uint32_t S[];
    constexpr uint32_t tile_words = kStride * WC + 4u;
    constexpr uint32_t group_words = TILES * 32u * WC;            // raw input of one group
    uint32_t *raw = S;                                            // NBUF * group_words
    uint32_t *slices = S + NBUF * group_words;                    // TILES * tile_words
    uint64_t *bar = reinterpret_cast<uint64_t *>(slices + TILES * tile_words);   // NBUF mbarriers (8-byte aligned)
    const uint32_t g = threadIdx.x / WC, c = threadIdx.x - g * WC;
    uint32_t *tile = slices + g * tile_words;
    const uint32_t tile_addr = smem_u32(tile);
    uint32_t addr[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) addr[j] = tile_addr + __ldg(slice_map + (uint32_t)j * WC + c);
    if (c == 0) tile[kStride * WC] = 0u;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NBUF; ++i) mbar_init(bar + i, 1);
        mbar_init_fence();
    }
    __syncthreads();
    pdl_enter();

    // bytes of group `grp` that lie inside the ciphertext (the last group may be ragged)
    auto issue = [&](uint64_t grp, uint32_t buf) {
        const uint64_t first_blk = grp * TILES * 32u;
        const uint64_t blocks = min((uint64_t)TILES * 32u, T - first_blk);
        const uint32_t bytes = (uint32_t)blocks * WC * 4u;
        mbar_expect_tx(bar + buf, bytes);
        bulk_g2s(raw + buf * group_words, in + first_blk * WC, bytes, bar + buf);
    };
    if (threadIdx.x == 0 && blockIdx.x < n_groups) issue(blockIdx.x, 0);

    uint32_t it = 0;
    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
        const uint32_t buf = NBUF == 2 ? (it & 1u) : 0u;
        const uint32_t parity = NBUF == 2 ? ((it >> 1) & 1u) : (it & 1u);
        const uint64_t nxt = grp + gridDim.x;
        if (NBUF == 2 && threadIdx.x == 0 && nxt < n_groups) {
            proxy_fence_async();             // the other buffer was read (phase A, previous iteration) before the last barrier
            issue(nxt, buf ^ 1u);
        }
        mbar_wait(bar + buf, parity);
        const uint64_t blk0 = (grp * TILES + g) * 32u;
        const bool full = blk0 + 32u <= T;
        {
            const uint32_t *col = raw + buf * group_words + g * 32u * WC + c;
            uint32_t x[32];
            if (full) {
#pragma unroll
                for (int b = 0; b < 32; ++b) x[b] = col[b * WC];
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b) x[b] = (blk0 + b < T) ? col[b * WC] : 0u;   // rows past the end were not copied
            }
            transpose32(x);
            store_slices(tile + kStride * c, x);
        }
        __syncthreads();
        if (NBUF == 1 && threadIdx.x == 0 && nxt < n_groups) {
            proxy_fence_async();             // every thread has read its column of the buffer before the barrier
            issue(nxt, 0);
        }
        uint32_t *dst = out + blk0 * WC + c;
        if (full) {
            uint32_t y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = lds_u32(addr[j]);
            column_out<WC, true>(y, dst, blk0, T, WC);
        } else {
            column_out_ragged<WC>(tile_addr, slice_map + c, dst, blk0, T, WC);
        }
        __syncthreads();   // before the next group overwrites the slices
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Plane kernel, for long blocks (W >= 64 32-bit words).  A tile of 32 blocks is 32 rows of W words; its 32*W slices
// are kept in the SAME 32 x W array -- slice (column c, bit j) at word j*W + c, "plane j" -- so a tile costs 128*W
// bytes of shared memory and nothing else (N = 16383: 64 KB; the padded slice rows of the kernels above need 74 KB
// next to a 64 KB raw tile, which leaves one CTA of 512 threads per SM).  Three CTAs of 256 threads are then
// resident per SM at N = 16383 and run out of phase: while one transposes (integer pipe), another gathers (shared
// memory) and the third loads or stores (HBM) -- the three resources this kernel needs in nearly equal measure.
//   BULK = true : the tile arrives by ONE bulk asynchronous copy (cp.async.bulk, byte-counted on an mbarrier) issued
//                 as soon as the previous tile's gathers are done; phase A transposes each column in place.
//   BULK = false: phase A loads its column from global memory into registers (32 coalesced 128-byte warp rows).
// Thread t owns columns t, t+THREADS, ...; lanes are consecutive columns, so every row access is conflict-free and
// only the gather (random planes and columns) sees bank conflicts.  plane_map[j*W + c] = BYTE offset (j_src*W +
// c_src)*4 of the source slice of output (column c, bit j), or of the zero words behind the tile for pad bits.
// (Keeping the thread's 32 gather offsets in registers, packed two to a word, was measured too: no faster -- the
// map loads are not what the kernel waits for -- and it spills at three CTAs per SM.)
template <int WC, int THREADS, int MINB, bool BULK>
__global__ void __launch_bounds__(THREADS, MINB)
permute_plane_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t Wrt,
                     const uint32_t *__restrict__ plane_map, uint32_t *__restrict__ out, const uint64_t n_tiles) {
    extern __shared__ __align__(128) uint32_t S[];
    const uint32_t W = WC ? (uint32_t)WC : Wrt;
    uint32_t *tile = S;                                                   // 32*W words, 4 zero words, the mbarrier
    uint64_t *bar = reinterpret_cast<uint64_t *>(S + 32u * W + 4u);
    const uint32_t tile_addr = smem_u32(tile);
    if (threadIdx.x < 4) tile[32u * W + threadIdx.x] = 0u;
    if (BULK && threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    pdl_enter();

    // a tile whose byte count is not a multiple of 16 (the ragged last tile of an odd-L ciphertext) is loaded by the
    // threads themselves; it can only be the last tile, so the mbarrier's phase stays in step with `it`
    auto bulk_bytes = [&](uint64_t t) -> uint32_t {
        const uint32_t blocks = (uint32_t)min((uint64_t)32u, T - t * 32u);
        const uint32_t bytes = blocks * W * 4u;
        return (bytes & 15u) ? 0u : bytes;
    };
    if (BULK && threadIdx.x == 0 && blockIdx.x < n_tiles) {
        const uint32_t bytes = bulk_bytes(blockIdx.x);
        if (bytes) {
            mbar_expect_tx(bar, bytes);
            bulk_g2s(tile, in + (uint64_t)blockIdx.x * 32u * W, bytes, bar);
        }
    }

    uint32_t it = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const uint64_t blk0 = t * 32u;
        const uint32_t blocks = (uint32_t)min((uint64_t)32u, T - blk0);
        const bool full = blocks == 32u;
        // ---- phase A: every column of the tile transposed into its 32 planes
        if (BULK) {
            if (bulk_bytes(t)) mbar_wait(bar, it & 1u);
            else {
                const uint32_t *src = in + blk0 * W;
                for (uint32_t idx = threadIdx.x; idx < blocks * W; idx += THREADS) tile[idx] = __ldcs(src + idx);
                __syncthreads();
            }
            for (uint32_t c = threadIdx.x; c < W; c += THREADS) {
                uint32_t x[32];
#pragma unroll
                for (int b = 0; b < 32; ++b) x[b] = tile[(uint32_t)b * W + c];   // rows past the end: stale words, whose
                transpose32(x);                                                  // bits only reach rows that are not stored
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[(uint32_t)j * W + c] = x[j];
            }
        } else {
            for (uint32_t c = threadIdx.x; c < W; c += THREADS) {
                const uint32_t *src = in + blk0 * W + c;
                uint32_t x[32];
                if (full) {
#pragma unroll
                    for (int b = 0; b < 32; ++b) x[b] = __ldcs(src + (uint32_t)b * W);
                } else {
#pragma unroll
                    for (int b = 0; b < 32; ++b) x[b] = ((uint32_t)b < blocks) ? __ldcs(src + (uint64_t)b * W) : 0u;
                }
                transpose32(x);
#pragma unroll
                for (int j = 0; j < 32; ++j) tile[(uint32_t)j * W + c] = x[j];
            }
        }
        __syncthreads();
        // ---- phase B: gather the 32 source planes of every output column, transpose back, store
        for (uint32_t c = threadIdx.x; c < W; c += THREADS) {
            const uint32_t *map = plane_map + c;
            uint32_t y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = lds_u32(tile_addr + __ldg(map + (uint32_t)j * W));
            transpose32(y);
            uint32_t *dst = out + blk0 * W + c;
            if (full) {
#pragma unroll
                for (int b = 0; b < 32; ++b) __stcs(dst + (uint32_t)b * W, y[b]);
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b)
                    if ((uint32_t)b < blocks) __stcs(dst + (uint64_t)b * W, y[b]);
            }
        }
        __syncthreads();   // every gather is done: the planes may be overwritten
        if (BULK && threadIdx.x == 0) {
            const uint64_t nxt = t + gridDim.x;
            if (nxt < n_tiles) {
                const uint32_t bytes = bulk_bytes(nxt);
                if (bytes) {
                    proxy_fence_async();
                    mbar_expect_tx(bar, bytes);
                    bulk_g2s(tile, in + nxt * 32u * W, bytes, bar);
                }
            }
        }
    }
}

// The plane layout for SHORT blocks (N = 1247: W = 40): a CTA works on a group of G tiles at once, one (tile, column)
// item per thread, each tile with its own zero words behind it (the raw rows of a tile arrive by one bulk copy per
// tile, all counted on one mbarrier).  20.5 KB of shared memory per CTA at W = 40, G = 4 instead of the 43.5 KB of the
// padded-slice kernel above, no hoisted addresses (56 registers instead of 96): six CTAs of 160 threads per SM.
// Measured and NOT adopted (instantiated only with -DCSGN_BUILD_VARIANTS): 55.8 us per 10^6 blocks at best against
// 54.2 us for permute_prefetch_kernel<40,4,1,4> -- the hoisted gather addresses and the 128-bit slice stores of that
// kernel are worth more than two extra resident CTAs (profiles/r2_perm_plane_group_sweep.log).
template <int WC, int G, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB)
permute_plane_group_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t *__restrict__ plane_map,
                           uint32_t *__restrict__ out, const uint64_t n_groups) {
    extern __shared__ __align__(128) uint32_t S[];
    constexpr uint32_t W = WC, tile_words = 32u * W + 4u;               // a tile and its 4 zero words
    uint64_t *bar = reinterpret_cast<uint64_t *>(S + G * tile_words);
    for (uint32_t i = threadIdx.x; i < 4u * G; i += THREADS) S[(i >> 2) * tile_words + 32u * W + (i & 3u)] = 0u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init_fence();
    }
    __syncthreads();
    pdl_enter();

    auto issue = [&](uint64_t grp) {
        const uint64_t first_blk = grp * (G * 32u);
        const uint32_t blocks = (uint32_t)min((uint64_t)(G * 32u), T - first_blk);
        mbar_expect_tx(bar, blocks * W * 4u);
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const uint32_t tb = blocks > 32u * g ? min(32u, blocks - 32u * g) : 0u;
            if (tb) bulk_g2s(S + g * tile_words, in + (first_blk + 32u * g) * W, tb * W * 4u, bar);
        }
    };
    if (threadIdx.x == 0 && blockIdx.x < n_groups) issue(blockIdx.x);

    uint32_t it = 0;
    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++it) {
        const uint64_t first_blk = grp * (G * 32u);
        const uint32_t blocks = (uint32_t)min((uint64_t)(G * 32u), T - first_blk);
        mbar_wait(bar, it & 1u);
        for (uint32_t item = threadIdx.x; item < G * W; item += THREADS) {
            const uint32_t g = item / W, c = item - g * W;
            uint32_t *tile = S + g * tile_words;
            uint32_t x[32];
#pragma unroll
            for (int b = 0; b < 32; ++b) x[b] = tile[(uint32_t)b * W + c];
            transpose32(x);
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[(uint32_t)j * W + c] = x[j];
        }
        __syncthreads();
        for (uint32_t item = threadIdx.x; item < G * W; item += THREADS) {
            const uint32_t g = item / W, c = item - g * W;
            const uint32_t tb = blocks > 32u * g ? min(32u, blocks - 32u * g) : 0u;   // rows of this tile inside the ciphertext
            if (tb == 0) continue;
            const uint32_t tile_addr = smem_u32(S + g * tile_words);
            uint32_t y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = lds_u32(tile_addr + __ldg(plane_map + (uint32_t)j * W + c));
            transpose32(y);
            uint32_t *dst = out + (first_blk + 32u * g) * W + c;
            if (tb == 32u) {
#pragma unroll
                for (int b = 0; b < 32; ++b) __stcs(dst + (uint32_t)b * W, y[b]);
            } else {
#pragma unroll
                for (int b = 0; b < 32; ++b)
                    if ((uint32_t)b < tb) __stcs(dst + (uint32_t)b * W, y[b]);
            }
        }
        __syncthreads();   // every gather is done: the planes may be overwritten
        if (threadIdx.x == 0) {
            const uint64_t nxt = grp + gridDim.x;
            if (nxt < n_groups) {
                proxy_fence_async();
                issue(nxt);
            }
        }
    }
}

// Any W (runtime): work items (tile, column) strided over the CTA's threads.
__global__ void __launch_bounds__(512, 2)
permute_sliced_kernel(const uint32_t *__restrict__ in, const uint64_t T, const uint32_t W,
                      const uint32_t *__restrict__ slice_map, uint32_t *__restrict__ out,
                      const uint32_t tiles_per_cta, const uint64_t n_groups) {
    extern __shared__ __align__(128) uint32_t S[];    // tiles_per_cta * (36*W + 4) words
    const uint32_t tile_words = kStride * W + 4u;
    const uint32_t items = tiles_per_cta * W;
    pdl_enter();

    for (uint32_t t = threadIdx.x; t < tiles_per_cta; t += blockDim.x) S[t * tile_words + kStride * W] = 0u;

    for (uint64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
        const uint64_t tile0 = grp * tiles_per_cta;
        for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
            const uint32_t tl = it / W, c = it - tl * W;
            const uint64_t blk0 = (tile0 + tl) * 32u;
            const uint32_t *src = in + blk0 * W + c;
            uint32_t *dst = S + tl * tile_words + kStride * c;
            if (blk0 + 32u <= T) column_in<0, true>(src, blk0, T, W, dst);
            else column_in_ragged<0>(src, blk0, T, W, dst);
        }
        __syncthreads();
        for (uint32_t it = threadIdx.x; it < items; it += blockDim.x) {
            const uint32_t tl = it / W, c = it - tl * W;
            const uint64_t blk0 = (tile0 + tl) * 32u;
            const uint32_t tile_addr = (uint32_t)__cvta_generic_to_shared(S + tl * tile_words);
            const uint32_t *map = slice_map + c;
            uint32_t *dst = out + blk0 * W + c;
            if (blk0 + 32u <= T) {
                uint32_t y[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = lds_u32(tile_addr + __ldg(map + (uint32_t)j * W));
                column_out<0, true>(y, dst, blk0, T, W);
            } else {
                column_out_ragged<0>(tile_addr, map, dst, blk0, T, W);
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

template <typename Kernel>
cudaError_t resident_ctas(Kernel kernel, int tpb, size_t smem, int *per_sm) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)device_props().smem_optin);
        if (e != cudaSuccess) return e;
    }
    int n = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, tpb, smem) != cudaSuccess || n < 1) n = 1;
    *per_sm = n;
    return cudaSuccess;
}

template <int WC, int TILES, bool HOIST, int MINB, int WAVES>
cudaError_t launch_fixed(const uint64_t *in, uint64_t T, const uint32_t *slice_map, uint64_t *out, cudaStream_t stream) {
    constexpr size_t smem = (size_t)TILES * (kStride * WC + 4u) * sizeof(uint32_t);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = resident_ctas(permute_fixed_kernel<WC, TILES, HOIST, MINB>, WC * TILES, smem, &per_sm);
        if (e != cudaSuccess) return e;
    }
    const uint64_t n_tiles = (T + 31) / 32;
    const uint64_t n_groups = (n_tiles + TILES - 1) / TILES;
    // WAVES x the resident CTAs loop over the groups: a few waves balance the tail better than a strictly
    // persistent grid (B200 sweeps in profiles/), while the per-CTA set-up stays amortised
    const uint64_t cap =
        (uint64_t)device_props().sm_count * per_sm * (uint64_t)std::max<long>(1, env_long("CSGN_PERM_WAVES", WAVES));
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_groups, cap));
    return launch_kernel(permute_fixed_kernel<WC, TILES, HOIST, MINB>, grid, WC * TILES, smem, stream,
                         reinterpret_cast<const uint32_t *>(in), T, slice_map, reinterpret_cast<uint32_t *>(out), n_groups);
}

template <int WC, int TILES, int NBUF, int MINB, int WAVES>
cudaError_t launch_prefetch(const uint64_t *in, uint64_t T, const uint32_t *slice_map, uint64_t *out, cudaStream_t stream) {
    constexpr size_t smem = ((size_t)NBUF * TILES * 32u * WC + (size_t)TILES * (kStride * WC + 4u)) * sizeof(uint32_t) +
                            NBUF * sizeof(uint64_t);
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = resident_ctas(permute_prefetch_kernel<WC, TILES, NBUF, MINB>, WC * TILES, smem, &per_sm);
        if (e != cudaSuccess) return e;
    }
    const uint64_t n_tiles = (T + 31) / 32;
    const uint64_t n_groups = (n_tiles + TILES - 1) / TILES;
    const uint64_t cap =
        (uint64_t)device_props().sm_count * per_sm * (uint64_t)std::max<long>(1, env_long("CSGN_PERM_WAVES", WAVES));
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_groups, cap));
    return launch_kernel(permute_prefetch_kernel<WC, TILES, NBUF, MINB>, grid, WC * TILES, smem, stream,
                         reinterpret_cast<const uint32_t *>(in), T, slice_map, reinterpret_cast<uint32_t *>(out), n_groups);
}

template <int WC, int THREADS, int MINB, bool BULK>
cudaError_t launch_plane(const uint64_t *in, uint64_t T, uint32_t W, const uint32_t *plane_map, uint64_t *out, int waves,
                         cudaStream_t stream) {
    const size_t smem = ((size_t)32 * W + 4) * sizeof(uint32_t) + 16;
    static int per_sm = 0;
    static size_t cached_smem = 0;
    if (per_sm == 0 || cached_smem != smem) {
        cudaError_t e = resident_ctas(permute_plane_kernel<WC, THREADS, MINB, BULK>, THREADS, smem, &per_sm);
        if (e != cudaSuccess) return e;
        cached_smem = smem;
    }
    const uint64_t n_tiles = (T + 31) / 32;
    const uint64_t cap = (uint64_t)device_props().sm_count * per_sm * (uint64_t)std::max<long>(1, env_long("CSGN_PERM_WAVES", waves));
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_tiles, cap));
    return launch_kernel(permute_plane_kernel<WC, THREADS, MINB, BULK>, grid, THREADS, smem, stream,
                         reinterpret_cast<const uint32_t *>(in), T, W, plane_map, reinterpret_cast<uint32_t *>(out), n_tiles);
}

template <int WC, int G, int THREADS, int MINB>
cudaError_t launch_plane_group(const uint64_t *in, uint64_t T, const uint32_t *plane_map, uint64_t *out, int waves,
                               cudaStream_t stream) {
    constexpr size_t smem = (size_t)G * (32u * WC + 4u) * sizeof(uint32_t) + 16;
    static int per_sm = 0;
    if (per_sm == 0) {
        cudaError_t e = resident_ctas(permute_plane_group_kernel<WC, G, THREADS, MINB>, THREADS, smem, &per_sm);
        if (e != cudaSuccess) return e;
    }
    const uint64_t n_groups = ((T + 31) / 32 + G - 1) / G;
    const uint64_t cap = (uint64_t)device_props().sm_count * per_sm * (uint64_t)std::max<long>(1, env_long("CSGN_PERM_WAVES", waves));
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_groups, cap));
    return launch_kernel(permute_plane_group_kernel<WC, G, THREADS, MINB>, grid, THREADS, smem, stream,
                         reinterpret_cast<const uint32_t *>(in), T, plane_map, reinterpret_cast<uint32_t *>(out), n_groups);
}

cudaError_t launch_sliced(const uint64_t *in, uint64_t T, uint32_t W, const uint32_t *slice_map, uint64_t *out,
                          uint32_t tiles_per_cta, uint32_t tpb, size_t smem, cudaStream_t stream) {
    const DeviceProps &dp = device_props();
    static int per_sm_cache = 0;
    static uint32_t cache_tpb = 0;
    static size_t cache_smem = 0;
    if (per_sm_cache == 0 || cache_tpb != tpb || cache_smem != smem) {
        cudaError_t e = resident_ctas(permute_sliced_kernel, (int)tpb, smem, &per_sm_cache);
        if (e != cudaSuccess) return e;
        cache_tpb = tpb;
        cache_smem = smem;
    }
    const uint64_t n_tiles = (T + 31) / 32;
    const uint64_t n_groups = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>(n_groups, (uint64_t)dp.sm_count * per_sm_cache * (uint64_t)env_long("CSGN_PERM_WAVES", 16)));
    return launch_kernel(permute_sliced_kernel, grid, tpb, smem, stream, reinterpret_cast<const uint32_t *>(in), T, W,
                         slice_map, reinterpret_cast<uint32_t *>(out), tiles_per_cta, n_groups);
}

}  // namespace

bool permute_sliced_supported(uint32_t L) {
    const size_t need = ((size_t)kStride * 2 * L + 4) * sizeof(uint32_t);
    return need <= device_props().smem_optin && 2 * L >= 1;
}

bool permute_plane_supported(uint32_t L) {
    return ((size_t)64 * L + 4) * sizeof(uint32_t) + 16 <= device_props().smem_optin;
}

cudaError_t launch_permute(const uint64_t *in, uint64_t T, uint32_t L, uint32_t N, const uint32_t *src_map,
                           const uint32_t *slice_map, const uint32_t *plane_map, uint64_t *out, cudaStream_t stream) {
    if (T == 0 || L == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const bool aligned4 = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3u) == 0;
    if (plane_map && aligned4 && permute_plane_supported(L) && !env_long("CSGN_PERM_GATHER", 0) &&
        !env_long("CSGN_PERM_VARIANT", 0)) {
        const uint32_t W = 2 * L;
        const long pv = env_long("CSGN_PERM_PLANE", -1);      // sweep knob: which instantiation
        const bool aligned16 = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;
        // The shipped forms (B200 sweeps: tools/perm_plane_sweep.py, profiles/r2_perm_plane_sweep*.log):
        //   W = 512 (N = 16383): 3 CTAs of 256 threads per SM, 4 waves   69 us / 90,000 blocks (was 75), 0.90 of the copy peak at 10^6
        //   W = 256 (N = 8191) : 6 CTAs of 128 threads per SM            57 us / 160,000 blocks (was 66)
        //   other 128 <= W <= 448: the runtime-W form, 6 CTAs of 128 threads
        // pv selects one of them by hand (tests); the other instantiations of the sweeps need -DCSGN_BUILD_VARIANTS.
        cudaError_t r = cudaErrorNotSupported;
        switch (pv) {
            case -1:
                if (W == 512 && aligned16) r = launch_plane<512, 256, 3, true>(in, T, W, plane_map, out, 4, stream);
                else if (W == 256 && aligned16) r = launch_plane<256, 128, 6, true>(in, T, W, plane_map, out, 4, stream);
                else if (W >= 128 && W <= 448 && aligned16) r = launch_plane<0, 128, 6, true>(in, T, W, plane_map, out, 16, stream);
                // (CTA sizes that waste fewer threads per column pass -- 192 for W = 130, 256 ... -- were measured:
                //  six small CTAs per SM beat them on four of five contexts, tools/perm_cta_probe.py)
                break;
            case 0: if (W == 512 && aligned16) r = launch_plane<512, 256, 3, true>(in, T, W, plane_map, out, 4, stream); break;
            case 6: r = launch_plane<0, 256, 3, false>(in, T, W, plane_map, out, 4, stream); break;
            case 12: if (W == 256 && aligned16) r = launch_plane<256, 128, 6, true>(in, T, W, plane_map, out, 4, stream); break;
            case 17: if (aligned16) r = launch_plane<0, 128, 6, true>(in, T, W, plane_map, out, 16, stream); break;
            case 18: if (aligned16) r = launch_plane<0, 192, 4, true>(in, T, W, plane_map, out, 16, stream); break;
#ifdef CSGN_BUILD_VARIANTS
            case 20: if (W == 40 && aligned16) r = launch_plane_group<40, 4, 160, 6>(in, T, plane_map, out, 2, stream); break;
            case 21: if (W == 40 && aligned16) r = launch_plane_group<40, 4, 160, 4>(in, T, plane_map, out, 2, stream); break;
            case 22: if (W == 40 && aligned16) r = launch_plane_group<40, 8, 320, 3>(in, T, plane_map, out, 2, stream); break;
            case 23: if (W == 40 && aligned16) r = launch_plane_group<40, 2, 96, 10>(in, T, plane_map, out, 2, stream); break;
            case 24: if (W == 40 && aligned16) r = launch_plane_group<40, 4, 96, 8>(in, T, plane_map, out, 2, stream); break;
            case 25: if (W == 40 && aligned16) r = launch_plane_group<40, 6, 256, 4>(in, T, plane_map, out, 2, stream); break;
            case 7: if (aligned16) r = launch_plane<0, 256, 3, true>(in, T, W, plane_map, out, 16, stream); break;
            case 1: if (W == 512) r = launch_plane<512, 256, 3, false>(in, T, W, plane_map, out, 1, stream); break;
            case 2: if (W == 512 && aligned16) r = launch_plane<512, 512, 1, true>(in, T, W, plane_map, out, 1, stream); break;
            case 3: if (W == 512) r = launch_plane<512, 512, 2, false>(in, T, W, plane_map, out, 1, stream); break;
            case 4: if (W == 512) r = launch_plane<512, 128, 6, false>(in, T, W, plane_map, out, 1, stream); break;
            case 5: if (W == 512 && aligned16) r = launch_plane<512, 128, 6, true>(in, T, W, plane_map, out, 1, stream); break;
            case 8: r = launch_plane<0, 128, 4, false>(in, T, W, plane_map, out, 1, stream); break;
            case 9: if (aligned16) r = launch_plane<0, 512, 1, true>(in, T, W, plane_map, out, 1, stream); break;
            case 10: if (W == 512 && aligned16) r = launch_plane<512, 512, 2, true>(in, T, W, plane_map, out, 1, stream); break;
            case 11: if (W == 512 && aligned16) r = launch_plane<512, 384, 2, true>(in, T, W, plane_map, out, 1, stream); break;
            case 13: if (W == 256 && aligned16) r = launch_plane<256, 256, 3, true>(in, T, W, plane_map, out, 1, stream); break;
            case 14: if (W == 256 && aligned16) r = launch_plane<256, 256, 4, true>(in, T, W, plane_map, out, 1, stream); break;
            case 15: if (W == 256) r = launch_plane<256, 128, 6, false>(in, T, W, plane_map, out, 1, stream); break;
            case 16: if (aligned16) r = launch_plane<0, 1024, 1, true>(in, T, W, plane_map, out, 1, stream); break;
#endif
            default: break;
        }
        if (r != cudaErrorNotSupported) {
            count_launch();
            return r;
        }
    }
    if (slice_map && aligned4 && permute_sliced_supported(L) && !env_long("CSGN_PERM_GATHER", 0)) {
        const uint32_t W = 2 * L;
        const long variant = env_long("CSGN_PERM_VARIANT", 0);   // 1: the runtime-W kernel even for known shapes
        count_launch();
        const bool aligned16 = (reinterpret_cast<uintptr_t>(in) & 15u) == 0;      // bulk copies need it
        if (W == 40 && variant == 0 && aligned16) {                                                            // N=1247
            // B200 sweeps (profiles/): 2 waves of resident CTAs up to a few million blocks, 4 beyond
            if (T >= 4000000) return launch_prefetch<40, 4, 1, 4, 4>(in, T, slice_map, out, stream);
            return launch_prefetch<40, 4, 1, 4, 2>(in, T, slice_map, out, stream);
        }
        if (W == 40 && (variant == 0 || variant == 8)) return launch_fixed<40, 4, true, 4, 4>(in, T, slice_map, out, stream);
#ifdef CSGN_BUILD_VARIANTS
        if (W == 40 && variant == 4 && aligned16) return launch_prefetch<40, 4, 2, 3, 1>(in, T, slice_map, out, stream);
        if (W == 40 && variant == 5 && aligned16) return launch_prefetch<40, 4, 1, 4, 1>(in, T, slice_map, out, stream);
        if (W == 40 && variant == 6 && aligned16) return launch_prefetch<40, 4, 2, 3, 4>(in, T, slice_map, out, stream);
        if (W == 40 && variant == 7 && aligned16) return launch_prefetch<40, 4, 1, 4, 4>(in, T, slice_map, out, stream);
        if (W == 40 && variant == 2) return launch_fixed<40, 4, false, 6, 4>(in, T, slice_map, out, stream);
        if (W == 40 && variant == 3) return launch_fixed<40, 8, true, 2, 4>(in, T, slice_map, out, stream);
        // N=16383 before the plane kernel: one 64 KB tile + 74 KB of slices = one CTA per SM (74.9 us / 90,000 blocks)
        if (W == 512 && variant == 9 && aligned16) return launch_prefetch<512, 1, 1, 1, 2>(in, T, slice_map, out, stream);
        if (W == 512 && variant == 8) return launch_fixed<512, 1, false, 2, 16>(in, T, slice_map, out, stream);
        if (W == 512 && variant == 2) return launch_fixed<512, 1, true, 1, 16>(in, T, slice_map, out, stream);
        if (W == 512 && variant == 4 && aligned16) return launch_prefetch<512, 1, 1, 1, 4>(in, T, slice_map, out, stream);
        if (W == 512 && variant == 5 && aligned16) return launch_prefetch<512, 1, 1, 1, 16>(in, T, slice_map, out, stream);
#endif
        const uint32_t tile_words = kStride * W + 4u;
        // several tiles per CTA when a block is short
        uint32_t tiles_per_cta = std::max<uint32_t>(1, (uint32_t)env_long("CSGN_PERM_ITEMS", 160) / W);
        const uint64_t n_tiles = (T + 31) / 32;
        tiles_per_cta = (uint32_t)std::min<uint64_t>(tiles_per_cta, n_tiles);
        while (tiles_per_cta > 1 && (size_t)tiles_per_cta * tile_words * 4 > 48 * 1024) --tiles_per_cta;
        const size_t smem = (size_t)tiles_per_cta * tile_words * sizeof(uint32_t);
        const uint32_t items = tiles_per_cta * W;
        const uint32_t tpb = std::min<uint32_t>(512, (items + 31) / 32 * 32);
        return launch_sliced(in, T, W, slice_map, out, tiles_per_cta, tpb, smem, stream);
    }
    const uint64_t total = T * L;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(
        1, std::min<uint64_t>((total + 255) / 256, (uint64_t)dp.sm_count * 8));
    count_launch();
    return launch_kernel(permute_gather_kernel, grid, 256, 0, stream, in, total, L, N, src_map, out);
}

}  // namespace csgn
