// mul.cu -- K1, ciphertext multiply: the all-pairs AND of T1 x T2 blocks -- and the fused multiply -> fold.
//
//   out[(i*T2+j)*L+k] = a[i*L+k] & b[j*L+k]     (reference src/Ciphertext.cpp:153-163;
//                                                the 1x1 shortcut :124-131 is T1=T2=1)
//
// Roofline: HBM WRITE bandwidth.  8*L bytes are written per output block, the
// operands (8*L*(T1+T2) bytes in total) stay L2-resident, one 64-bit AND per 8 bytes.
//
// A ciphertext is a flat stream of units: 16 bytes (uint4, UPB = L/2 units per block) when L is even and
// the words are 16-byte aligned, 8 bytes (uint2, UPB = L) otherwise (odd L: N = 191, 4097 ...).
//
// Tiled kernel (mul_outer_kernel).  Output row i is the whole right operand, seen as one flat
// stream of Q = T2*UPB units, ANDed with block a_i repeated with period UPB.  The CTA size is a multiple
// of UPB, so a thread that walks the flat stream with stride blockDim always meets the same fragment of a_i:
//   - a work item is (column tile of U*blockDim units) x (chunk of R rows);
//   - the thread's U units of b are loaded ONCE into registers (coalesced);
//   - the R blocks of a are staged in shared memory;
//   - per row: one shared load for the thread's fragment of a_i, U ANDs, U coalesced
//     streaming stores (st.global.cs -- nothing re-reads the product).
// b is re-read from L2 once per R rows, so L2 read traffic is 1/R of the write
// stream; no integer division happens inside the row loop.
//
// Chains ((a*b)*d with a fresh, short d: 10^6 rows of a few thousand units) use the same kernel with few rows per
// item, so that the resident CTAs write one compact, advancing window of the product (7.1 TB/s on a 20 GB product).
// A flat kernel (b resident in shared memory, the output walked as ONE stream in grid-stride steps, a fragments
// fetched a step ahead: mul_flat_kernel, built only with -DCSGN_BUILD_VARIANTS) was measured against it on both chain
// shapes and never won (profiles/r2_sweep_chain.log): 6.4-6.5 TB/s at best.
//
// Fused multiply -> fold (FOLD != 0): SecretKey::decrypt of the product (reference src/SecretKey.cpp:131-140,
// the caller pattern tests/basic_operations.cpp:35-40) evaluated on the product units while they are still in
// registers: the thread tests a_i & b_j against ITS unit of the key mask (its unit index within a block never
// changes), the per-unit verdicts of a tile are OR-ed across the UPB adjacent threads that hold one block, and the
// satisfied blocks are counted and folded into the grid total exactly as in decrypt.cu.  One pass over HBM (the
// product is written, never read back); FOLD == 2 stores nothing at all (decrypt-only consumers).
#include "kernels.cuh"
#include "launch.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cstring>

namespace csgn {
namespace {

constexpr int kMulMaxThreads = 512;
constexpr uint32_t kMulMaxSmem = 32 * 1024;
#ifdef CSGN_BUILD_VARIANTS
constexpr uint32_t kFlatMaxSmem = 64 * 1024;     // right operand resident in shared memory
constexpr bool kFlatByDefault = false;           // never: the tiled kernel with few rows per item wins on every chain shape
#endif

__device__ __forceinline__ uint4 vand(const uint4 a, const uint4 b) {
    return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w);
}
__device__ __forceinline__ uint2 vand(const uint2 a, const uint2 b) { return make_uint2(a.x & b.x, a.y & b.y); }
template <typename VT> __device__ __forceinline__ VT vzero();
template <> __device__ __forceinline__ uint4 vzero<uint4>() { return make_uint4(0u, 0u, 0u, 0u); }
template <> __device__ __forceinline__ uint2 vzero<uint2>() { return make_uint2(0u, 0u); }

// What a fused launch carries besides the operands (by value in the kernel parameters).
struct FoldParams {
    ParamMask pmask;            // the key mask, when it fits (always hot: parameter bank)
    const void *mask;           // the key mask in global memory (any size)
    uint64_t *scratch;          // the launch's fold scratch word
    uint64_t *count_out;        // device word for the total (may be null with a PeerPush)
    PeerPush pp;
    int mask_in_params;
    // Odd L multiplied with 16-byte units (plain multiply, even T2): the right operand is read as T2/2 double blocks of
    // 2L words, and each row of A -- dbl_words = L 8-byte words in memory -- is staged as the double block a_i || a_i.
    uint32_t dbl_words;
};

template <typename VT>
__device__ __forceinline__ VT fold_mask_unit(const FoldParams &fo, const uint32_t k) {
    return fo.mask_in_params ? reinterpret_cast<const VT *>(&fo.pmask)[k] : __ldg(static_cast<const VT *>(fo.mask) + k);
}

// OR of `fails` over the UPB adjacent threads that hold one block; the group's first thread gets the result and
// counts the clear bits under `valid`.  Whole CTA.  sFail: blockDim.x words of shared memory.
__device__ __forceinline__ uint32_t count_group_clear(const uint64_t fails, const uint64_t valid, const uint32_t UPB,
                                                      const uint32_t k, uint64_t *sFail) {
    uint32_t got = 0;
    if ((UPB & 31u) == 0u) {
        // groups are whole warps (the CTA size is a multiple of UPB, hence of 32): reduce in the warp first
        const uint32_t lo = __reduce_or_sync(0xffffffffu, (uint32_t)fails);
        const uint32_t hi = __reduce_or_sync(0xffffffffu, (uint32_t)(fails >> 32));
        if ((threadIdx.x & 31u) == 0u) sFail[threadIdx.x >> 5] = ((uint64_t)hi << 32) | lo;
        __syncthreads();
        if (k == 0) {
            uint64_t f = 0;
            const uint32_t w0 = threadIdx.x >> 5;
            for (uint32_t i = 0; i < (UPB >> 5); ++i) f |= sFail[w0 + i];
            got = (uint32_t)__popcll(~f & valid);
        }
    } else {
        sFail[threadIdx.x] = fails;
        __syncthreads();
        if (k == 0) {
            uint64_t f = 0;
            for (uint32_t i = 0; i < UPB; ++i) f |= sFail[threadIdx.x + i];
            got = (uint32_t)__popcll(~f & valid);
        }
    }
    return got;
}

// CTA count (held by the groups' first threads; almost always zero) -> grid.  Whole CTA, once.
__device__ __forceinline__ void publish_cta_count(const uint32_t cnt, unsigned long long *s_cnt, const FoldParams &fo) {
    if (cnt) atomicAdd(s_cnt, (unsigned long long)cnt);
    __syncthreads();
    grid_publish(threadIdx.x == 0 ? (uint64_t)*s_cnt : 0ull, fo.scratch, fo.count_out, fo.pp);
}

// Fail bit(s) of product unit `v` (unit u of the thread) under mask unit `m`: bit u, and for double blocks bit u for the
// first block of the pair, bit 32+u for the second.
template <typename VT, int DBLF>
__device__ __forceinline__ uint64_t fail_bits(const VT v, const VT m, const int u, const bool lo_first, const bool hi_first) {
    if constexpr (DBLF != 0 && sizeof(VT) == 16) {
        const bool fl = ((~v.x & m.x) | (~v.y & m.y)) != 0u, fh = ((~v.z & m.z) | (~v.w & m.w)) != 0u;
        const bool first = (lo_first && fl) || (hi_first && fh), second = (!lo_first && fl) || (!hi_first && fh);
        return (first ? (1ull << u) : 0ull) | (second ? (1ull << (32 + u)) : 0ull);
    } else {
        return unit_fails(v, m) ? (1ull << u) : 0ull;
    }
}

// FOLD: 0 = multiply, 1 = multiply and count the satisfied product blocks, 2 = count only (nothing stored).
// ALIGN (fused kernels, blocks of up to 16 units): a warp uses its first (32/UPB)*UPB lanes, so that the UPB threads
// holding one block are always lanes of ONE warp -- the per-block OR of the fail bits is then two redux.sync per item
// and needs no shared memory and no barrier.  (N = 1247: 30 of 32 lanes; the idle lanes cost issue slots of a kernel
// that is bound by HBM, not by issue.)
// DBLF (fused, odd L on 16-byte units): a "block" of UPB = L units is TWO blocks; a unit's low word belongs to the first
// when its word index 2k is below L, its high word when 2k+1 is.  The fail bits of the two halves travel in the low and
// the high 32 bits of the thread's fail word (rows x units <= 32), so one OR across the group still serves both.
// DBLA: the rows of A are staged doubled (a_i || a_i) -- set for every double-block launch, fused or not; a template
// argument so that the ordinary kernels carry none of it (the chain-shape multiply lost 7 % to a run-time test here).
template <typename VT, int U, int FOLD, int ALIGN, int DBLF = 0, int DBLA = 0>
__global__ void __launch_bounds__(kMulMaxThreads)
mul_outer_kernel(const VT *__restrict__ A, const VT *__restrict__ B, VT *__restrict__ out,
                 const uint32_t UPB, const uint64_t T1, const uint64_t Q, const uint32_t R,
                 const uint32_t n_col_tiles, const uint64_t n_items, const uint32_t pf_chunks,
                 const __grid_constant__ FoldParams fo) {
    extern __shared__ uint4 smem_raw[];
    VT *sA = reinterpret_cast<VT *>(smem_raw);
    // fused fold: one 64-bit fail word per thread, behind the staged rows (8-byte aligned: R*UPB units of 8 or 16 bytes)
    uint64_t *sFail = reinterpret_cast<uint64_t *>(sA + (size_t)R * UPB);
    __shared__ unsigned long long s_cnt;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t lanes_used = ALIGN ? (32u / UPB) * UPB : 32u;             // ALIGN: whole blocks per warp
    const bool active = !ALIGN || lane < lanes_used;
    // this thread's unit inside a row step, and the units one row step of the CTA covers
    const uint32_t my_unit = ALIGN ? (threadIdx.x >> 5) * lanes_used + lane : threadIdx.x;
    const uint32_t tpb = ALIGN ? (blockDim.x >> 5) * lanes_used : blockDim.x;
    const uint32_t k = my_unit % UPB;
    const uint32_t gmask = ALIGN ? ((UPB >= 32u ? 0xffffffffu : ((1u << UPB) - 1u)) << (lane / UPB * UPB)) : 0u;
    const uint64_t tile_q = (uint64_t)tpb * U;
    VT m = vzero<VT>();
    bool lo_first = true, hi_first = true;         // DBLF: which block of the pair this unit's words belong to
    if (FOLD) {
        if constexpr (DBLF != 0 && sizeof(VT) == 16) {
            const uint32_t Lw = fo.dbl_words;
            const uint64_t *M64 = static_cast<const uint64_t *>(fo.mask);
            lo_first = 2u * k < Lw;
            hi_first = 2u * k + 1u < Lw;
            const uint64_t mlo = __ldg(M64 + (lo_first ? 2u * k : 2u * k - Lw));
            const uint64_t mhi = __ldg(M64 + (hi_first ? 2u * k + 1u : 2u * k + 1u - Lw));
            m = make_uint4((uint32_t)mlo, (uint32_t)(mlo >> 32), (uint32_t)mhi, (uint32_t)(mhi >> 32));
        } else {
            m = fold_mask_unit<VT>(fo, k);         // the key is not the predecessor's output: before the PDL wait
        }
        if (threadIdx.x == 0) s_cnt = 0ull;
    }
    pdl_enter();

    // In real use the left operand is DRAM-cold (the previous product flushed L2) and an
    // item is short, so a cold a-chunk is a full DRAM latency in front of every item.  CTAs
    // start in blockIdx order, so CTA j warms L2 for the items that start ~two waves later:
    // the chunk of item j + pf_chunks*n_col_tiles (done by the column-0 CTA of each chunk);
    // the first pf_chunks chunks, which nobody is ahead of, are requested line by line by
    // the first CTAs.
    const uint32_t a_row_bytes = DBLA ? fo.dbl_words * 8u : UPB * (uint32_t)sizeof(VT);   // a row of A in memory
    if (pf_chunks) {
        const uint64_t a_bytes = T1 * a_row_bytes;
        const uint64_t head_bytes = min(a_bytes, (uint64_t)pf_chunks * R * a_row_bytes);
        if (threadIdx.x == 0 && (uint64_t)blockIdx.x * 128u < head_bytes)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A) + (uint64_t)blockIdx.x * 128u));
        if (blockIdx.x % n_col_tiles == 0) {
            const uint64_t row0 = ((uint64_t)blockIdx.x / n_col_tiles + pf_chunks) * R;
            if (row0 < T1) {
                const uint32_t bytes = (uint32_t)min((uint64_t)R, T1 - row0) * a_row_bytes;
                const uint32_t off = threadIdx.x * 128u;
                if (off < bytes + 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A) + row0 * a_row_bytes +
                                                                   min(off, bytes - 1u)));
            }
        }
    }

    uint32_t cnt = 0;
    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t ct = (uint32_t)(item % n_col_tiles);
        const uint64_t row0 = (item / n_col_tiles) * R;
        const uint32_t nrows = (uint32_t)min((uint64_t)R, T1 - row0);
        const uint64_t q0 = (uint64_t)ct * tile_q + my_unit;

        VT b[U];
        uint32_t live_bits = 0;                     // bit u: unit u of this thread lies inside the row
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t q = q0 + (uint64_t)u * tpb;
            const bool in = active && q < Q;
            b[u] = in ? __ldg(B + q) : vzero<VT>();
            live_bits |= in ? (1u << u) : 0u;
        }

        __syncthreads();  // the previous item's readers of sA (and sFail) are done
        if constexpr (DBLA != 0 && sizeof(VT) == 16) {
            // a_i || a_i: unit u of the double block is words (2u mod L, (2u+1) mod L) of a_i
            const uint32_t Lw = fo.dbl_words;
            const uint64_t *a64 = reinterpret_cast<const uint64_t *>(A) + row0 * Lw;
            for (uint32_t idx = threadIdx.x; idx < nrows * UPB; idx += blockDim.x) {
                const uint32_t r = idx / UPB, u = idx - r * UPB;
                const uint32_t w0 = 2u * u < Lw ? 2u * u : 2u * u - Lw, w1 = 2u * u + 1u < Lw ? 2u * u + 1u : 2u * u + 1u - Lw;
                const uint64_t x0 = __ldg(a64 + (uint64_t)r * Lw + w0), x1 = __ldg(a64 + (uint64_t)r * Lw + w1);
                reinterpret_cast<uint2 *>(sA)[2 * idx] = make_uint2((uint32_t)x0, (uint32_t)(x0 >> 32));
                reinterpret_cast<uint2 *>(sA)[2 * idx + 1] = make_uint2((uint32_t)x1, (uint32_t)(x1 >> 32));
            }
        } else {
            const VT *a_chunk = A + row0 * UPB;
            for (uint32_t idx = threadIdx.x; idx < nrows * UPB; idx += blockDim.x) sA[idx] = __ldg(a_chunk + idx);
        }
        __syncthreads();

        VT *o = out + row0 * Q + q0;
        const VT *sa = sA + k;
        uint64_t fails = 0;                         // row r, unit u -> bit (nrows-1-r)*U + u
        if (ALIGN && !active) {
            // idle lanes of a lane-aligned warp: nothing to load, store or vote
        } else if ((uint64_t)(ct + 1) * tile_q <= Q) {
#pragma unroll 4
            for (uint32_t r = 0; r < nrows; ++r, o += Q, sa += UPB) {
                const VT a = *sa;
                uint64_t rowbits = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const VT v = vand(a, b[u]);
                    if (FOLD != 2) __stcs(o + (uint64_t)u * tpb, v);
                    if (FOLD) rowbits |= fail_bits<VT, DBLF>(v, m, u, lo_first, hi_first);
                }
                if (FOLD) fails = (fails << U) | rowbits;
            }
        } else {
            // ragged last column tile: per-unit bounds, loop-invariant predicates
            for (uint32_t r = 0; r < nrows; ++r, o += Q, sa += UPB) {
                const VT a = *sa;
                uint64_t rowbits = 0;
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const VT v = vand(a, b[u]);
                    if (FOLD != 2 && ((live_bits >> u) & 1u)) __stcs(o + (uint64_t)u * tpb, v);
                    if (FOLD) rowbits |= fail_bits<VT, DBLF>(v, m, u, lo_first, hi_first);
                }
                if (FOLD) fails = (fails << U) | rowbits;
            }
        }
        if (FOLD) {
            // nrows*U <= 64 (launcher); live units repeated for every row
            const uint32_t nb = nrows * U;
            const uint64_t all_rows = nb >= 64u ? ~0ull : ((1ull << nb) - 1ull);
            uint64_t valid = (all_rows / ((1ull << U) - 1ull)) * (uint64_t)live_bits;
            if (DBLF) valid |= valid << 32;              // both halves of every double block (nb <= 32, launcher)
            if (ALIGN) {
                if (active) {
                    uint64_t f = (uint64_t)__reduce_or_sync(gmask, (uint32_t)fails);
                    if (nb > 32u) f |= (uint64_t)__reduce_or_sync(gmask, (uint32_t)(fails >> 32)) << 32;   // CTA-uniform
                    if (k == 0) cnt += (uint32_t)__popcll(~f & valid);
                }
            } else {
                cnt += count_group_clear(fails, valid, UPB, k, sFail);
            }
        }
    }
    if (FOLD) publish_cta_count(cnt, &s_cnt, fo);
}

#ifdef CSGN_BUILD_VARIANTS
// Flat kernel: see the head of the file.  Requires UPB | blockDim, blockDim <= Q, Q units of b in shared memory.
// Step s of the CTA covers units [s*W, (s+1)*W) of the output stream, W = blockDim*U; steps are dealt round-robin
// over the grid.  (i, q) = (row, unit in row) of the thread's first unit advance by a precomputed (di, dq) per
// step -- no division in the loop; the thread's U units are blockDim apart and wrap at most once each.
template <typename VT, int U, int FOLD>
__global__ void __launch_bounds__(kMulMaxThreads)
mul_flat_kernel(const VT *__restrict__ A, const VT *__restrict__ B, VT *__restrict__ out, const uint32_t UPB,
                const uint32_t Q, const uint64_t total_units, const uint64_t n_steps, const uint64_t di,
                const uint32_t dq, const __grid_constant__ FoldParams fo) {
    extern __shared__ uint4 smem_raw[];
    VT *sB = reinterpret_cast<VT *>(smem_raw);
    uint64_t *sFail = reinterpret_cast<uint64_t *>(sB + Q);
    __shared__ unsigned long long s_cnt;
    const uint32_t tpb = blockDim.x;
    const uint32_t k = threadIdx.x % UPB;
    const uint64_t W = (uint64_t)tpb * U;
    VT m = vzero<VT>();
    if (FOLD) {
        m = fold_mask_unit<VT>(fo, k);
        if (threadIdx.x == 0) s_cnt = 0ull;
    }
    pdl_enter();
    for (uint32_t idx = threadIdx.x; idx < Q; idx += tpb) sB[idx] = __ldg(B + idx);

    // (row, unit in row) of this thread's first unit in step blockIdx.x
    uint64_t i;
    uint32_t q;
    {
        const uint64_t g0 = (uint64_t)blockIdx.x * W + threadIdx.x;
        i = g0 / Q;                                  // once per thread
        q = (uint32_t)(g0 - i * Q);
    }
    const VT *Ak = A + k;
    const uint64_t T1 = total_units / Q;

    // a fragments of the current step; the next step's are requested before the current one is stored
    VT a_cur[U];
    {
        uint64_t iu = i;
        uint32_t qu = q;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            a_cur[u] = iu < T1 ? __ldg(Ak + iu * UPB) : vzero<VT>();
            qu += tpb;
            if (qu >= Q) { qu -= Q; ++iu; }
        }
    }
    __syncthreads();

    uint32_t cnt = 0;
    uint64_t fails = 0;
    uint32_t nbits = 0;
    for (uint64_t s = blockIdx.x; s < n_steps; s += gridDim.x) {
        // next step's position and fragments
        uint64_t i_n = i + di;
        uint32_t q_n = q + dq;
        if (q_n >= Q) { q_n -= Q; ++i_n; }
        VT a_nxt[U];
        {
            uint64_t iu = i_n;
            uint32_t qu = q_n;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                a_nxt[u] = iu < T1 ? __ldg(Ak + iu * UPB) : vzero<VT>();
                qu += tpb;
                if (qu >= Q) { qu -= Q; ++iu; }
            }
        }
        VT *o = out + (s * W + threadIdx.x);
        uint64_t g = s * W + threadIdx.x;
        uint32_t qu = q;
        uint32_t stepbits = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool in = g < total_units;
            const VT v = vand(a_cur[u], sB[qu]);
            if (FOLD != 2 && in) __stcs(o, v);
            if (FOLD) stepbits |= (!in || unit_fails(v, m)) ? (1u << u) : 0u;
            o += tpb;
            g += tpb;
            qu += tpb;
            if (qu >= Q) qu -= Q;
        }
        if (FOLD) {
            fails = (fails << U) | stepbits;
            nbits += U;
            if (nbits + U > 64u) {                   // CTA-uniform: every thread runs the same steps
                cnt += count_group_clear(fails, nbits >= 64u ? ~0ull : ((1ull << nbits) - 1ull), UPB, k, sFail);
                __syncthreads();                     // sFail is reused by the next flush
                fails = 0;
                nbits = 0;
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) a_cur[u] = a_nxt[u];
        i = i_n;
        q = q_n;
    }
    if (FOLD) {
        if (nbits) cnt += count_group_clear(fails, (1ull << nbits) - 1ull, UPB, k, sFail);
        publish_cta_count(cnt, &s_cnt, fo);
    }
}

#endif  // CSGN_BUILD_VARIANTS

// Blocks longer than kMulMaxThreads units (N > 65536), any alignment: one 64-bit word per thread-iteration.
__global__ void __launch_bounds__(256)
mul_outer_generic_kernel(const uint64_t *__restrict__ A, const uint64_t *__restrict__ B,
                         uint64_t *__restrict__ out, const uint32_t L, const uint64_t row_words,
                         const uint64_t total_words) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    pdl_enter();
    for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_words; idx += stride) {
        const uint64_t i = idx / row_words;
        const uint64_t in_row = idx - i * row_words;
        const uint32_t k = (uint32_t)(in_row % L);
        __stcs(out + idx, __ldg(A + i * L + k) & __ldg(B + in_row));
    }
}

void fill_fold_params(FoldParams &fo, const MulFold *fold, uint32_t upb, size_t unit_bytes, uint32_t dbl_words = 0) {
    memset(&fo, 0, sizeof fo);
    fo.dbl_words = dbl_words;
    if (!fold) return;
    fo.mask = fold->mask;
    fo.scratch = fold->scratch;
    fo.count_out = fold->count_out;
    if (fold->peer) fo.pp = *fold->peer;
    if (dbl_words == 0 && fold->host_mask && (size_t)upb * unit_bytes <= sizeof(ParamMask)) {   // (double blocks read the L-word mask from global memory)
        memcpy(&fo.pmask, fold->host_mask, (size_t)upb * unit_bytes);
        fo.mask_in_params = 1;
    }
}

template <typename VT, int U, int FOLD, int ALIGN = 0, int DBLF = 0, int DBLA = 0>
cudaError_t launch_tiled_uf(const void *a, uint64_t T1, const void *b, uint64_t Q, uint32_t upb, void *out, uint32_t tpb,
                            uint32_t R, uint64_t grid_cap, const MulFold *fold, cudaStream_t stream, uint32_t dbl_words = 0) {
    // ALIGN: tpb threads (a multiple of 32) cover (tpb/32) * (32/upb)*upb units per row step
    const uint64_t tile_q = (uint64_t)(ALIGN ? (tpb / 32u) * ((32u / upb) * upb) : tpb) * U;
    const uint64_t n_col_tiles = (Q + tile_q - 1) / tile_q;
    const uint64_t n_chunks = (T1 + R - 1) / R;
    const uint64_t n_items = n_col_tiles * n_chunks;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, grid_cap);
    const size_t smem = (size_t)R * upb * sizeof(VT) + ((FOLD && !ALIGN) ? (size_t)tpb * sizeof(uint64_t) : 0);
    // prefetch distance: the chunks that ~two waves of resident CTAs cover
    const long pf_per_sm = env_long("CSGN_MUL_PF_CTAS_PER_SM", 8);
    const uint64_t ahead_items = (uint64_t)device_props().sm_count * (uint64_t)std::max<long>(pf_per_sm, 0);
    const uint32_t pf_chunks = pf_per_sm > 0 ? (uint32_t)((ahead_items + n_col_tiles - 1) / n_col_tiles) : 0u;
    FoldParams fo;
    fill_fold_params(fo, fold, upb, sizeof(VT), dbl_words);
    return launch_kernel(mul_outer_kernel<VT, U, FOLD, ALIGN, DBLF, DBLA>, grid, tpb, smem, stream, static_cast<const VT *>(a),
                         static_cast<const VT *>(b), static_cast<VT *>(out), upb, T1, Q, R, (uint32_t)n_col_tiles, n_items,
                         pf_chunks, fo);
}

template <typename VT, int U>
cudaError_t launch_tiled_u(int fold_mode, bool align, const void *a, uint64_t T1, const void *b, uint64_t Q, uint32_t upb,
                           void *out, uint32_t tpb, uint32_t R, uint64_t grid_cap, const MulFold *fold, cudaStream_t stream,
                           uint32_t dbl_words = 0) {
    if (align) {
        if (fold_mode == 1) return launch_tiled_uf<VT, U, 1, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream);
        return launch_tiled_uf<VT, U, 2, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream);
    }
    if constexpr (sizeof(VT) == 16 && U <= 2) {
        if (dbl_words && fold_mode == 1)
            return launch_tiled_uf<VT, U, 1, 0, 1, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream, dbl_words);
        if (dbl_words && fold_mode == 2)
            return launch_tiled_uf<VT, U, 2, 0, 1, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream, dbl_words);
    }
    if (dbl_words && fold_mode) return cudaErrorNotSupported;       // the launcher caps U at 2 for fused double blocks
    switch (fold_mode) {
        case 1: return launch_tiled_uf<VT, U, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream);
        case 2: return launch_tiled_uf<VT, U, 2>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream);
        default:
            if constexpr (sizeof(VT) == 16) {
                if (dbl_words) return launch_tiled_uf<VT, U, 0, 0, 0, 1>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream, dbl_words);
            }
            return launch_tiled_uf<VT, U, 0>(a, T1, b, Q, upb, out, tpb, R, grid_cap, fold, stream);
    }
}

#ifdef CSGN_BUILD_VARIANTS
template <typename VT, int U, int FOLD>
cudaError_t launch_flat_uf(const void *a, uint64_t T1, const void *b, uint32_t Q, uint32_t upb, void *out, uint32_t tpb,
                           uint32_t grid_cap, const MulFold *fold, cudaStream_t stream) {
    const uint64_t total = T1 * (uint64_t)Q;
    const uint64_t W = (uint64_t)tpb * U;
    const uint64_t n_steps = (total + W - 1) / W;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_steps, grid_cap);
    const uint64_t delta = (uint64_t)grid * W;
    const size_t smem = (size_t)Q * sizeof(VT) + (FOLD ? (size_t)tpb * sizeof(uint64_t) : 0);
    static size_t configured = 0;       // per instantiation
    if (smem > 48 * 1024 && smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(mul_flat_kernel<VT, U, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kFlatMaxSmem + (int)(kMulMaxThreads * sizeof(uint64_t)));
        if (e != cudaSuccess) return e;
        configured = kFlatMaxSmem + kMulMaxThreads * sizeof(uint64_t);
    }
    FoldParams fo;
    fill_fold_params(fo, fold, upb, sizeof(VT));
    return launch_kernel(mul_flat_kernel<VT, U, FOLD>, grid, tpb, smem, stream, static_cast<const VT *>(a),
                         static_cast<const VT *>(b), static_cast<VT *>(out), upb, Q, total, n_steps, delta / Q,
                         (uint32_t)(delta % Q), fo);
}

template <typename VT, int U>
cudaError_t launch_flat_u(int fold_mode, const void *a, uint64_t T1, const void *b, uint32_t Q, uint32_t upb, void *out,
                          uint32_t tpb, uint32_t grid_cap, const MulFold *fold, cudaStream_t stream) {
    switch (fold_mode) {
        case 1: return launch_flat_uf<VT, U, 1>(a, T1, b, Q, upb, out, tpb, grid_cap, fold, stream);
        case 2: return launch_flat_uf<VT, U, 2>(a, T1, b, Q, upb, out, tpb, grid_cap, fold, stream);
        default: return launch_flat_uf<VT, U, 0>(a, T1, b, Q, upb, out, tpb, grid_cap, fold, stream);
    }
}

#endif  // CSGN_BUILD_VARIANTS

// CTA size for the tiled kernel: a multiple of UPB in [192, cap].  A partial last warp costs
// issue slots in every item; a ragged last column tile only makes that tile's items shorter
// (items are scheduled dynamically), so it weighs a quarter.
uint32_t pick_tpb(uint32_t upb, uint64_t Q, int U, uint32_t cap) {
    uint32_t best = 0;
    double best_cost = 1e30;
    for (uint32_t t = (cap / upb) * upb; t >= upb && t > 0; t -= upb) {
        const uint64_t tile = (uint64_t)t * U;
        const uint64_t n_tiles = (Q + tile - 1) / tile;
        const double pad = (double)(n_tiles * tile) / (double)Q;
        const double warp = (double)((t + 31) / 32 * 32) / (double)t;
        const double cost = warp * (1.0 + 0.25 * (pad - 1.0)) * (t < 256 ? 1.03 : 1.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = t;
        }
        if (t < 192 + upb) break;
    }
    return best;
}

#ifdef CSGN_BUILD_VARIANTS
// The multiple of `upb` in [lo, hi] that wastes the fewest lanes of a partial warp (largest on ties).
uint32_t pick_tpb_flat(uint32_t upb, uint32_t lo, uint32_t hi) {
    uint32_t best = 0;
    double best_cost = 1e30;
    for (uint32_t t = (hi / upb) * upb; t >= upb && t >= lo; t -= upb) {
        const double warp = (double)((t + 31) / 32 * 32) / (double)t;
        if (warp < best_cost - 1e-9) {
            best_cost = warp;
            best = t;
        }
    }
    return best;
}

#endif  // CSGN_BUILD_VARIANTS

template <typename VT>
cudaError_t launch_units(int fold_mode, const void *a, uint64_t T1, const void *b, uint64_t T2, uint32_t upb, void *out,
                         const MulFold *fold, cudaStream_t stream, uint32_t dbl_words = 0) {
    const DeviceProps &dp = device_props();
    const uint64_t Q = T2 * upb;
    // The fused kernels hold more registers (54-64 against 34-40): CTAs of 256 threads keep four of them resident per
    // SM, which is what hides an item's load phase behind the other CTAs' stores (B200, tools/fused_tpb_sweep.py: 22.4 us
    // per 10^6-block product with 500-thread CTAs, 21.7 with 256 -- the plain multiply's time).
    const uint32_t tpb_cap = (uint32_t)std::min<long>(kMulMaxThreads, env_long("CSGN_MUL_TPB", fold_mode ? 256 : 512));
    const uint64_t out_units = T1 * Q;
    const bool huge = out_units * sizeof(VT) >= (1ull << 30);     // >= 1 GiB of output
    const uint64_t grid_cap = (uint64_t)std::min<long>(1l << 23, env_long("CSGN_MUL_GRID", 1l << 23));

#ifdef CSGN_BUILD_VARIANTS
    // ---- flat kernel: short right operand, very many rows (chain products) ----
    const long flat_knob = env_long("CSGN_MUL_FLAT", -1);         // -1: heuristic, 0: never, 1: whenever legal
    const bool flat_legal = Q * sizeof(VT) <= kFlatMaxSmem && Q >= upb && Q >= 64 && T1 >= 2;
    const bool flat_wanted = flat_knob > 0 || (flat_knob < 0 && kFlatByDefault && huge && T1 >= 16 * (uint64_t)dp.sm_count);
    if (flat_legal && flat_wanted && dbl_words == 0) {      // the flat kernel has no double-block staging
        int U = (int)env_long("CSGN_MUL_FLAT_U", 4);
        U = U >= 8 ? 8 : U >= 4 ? 4 : U >= 2 ? 2 : 1;
        uint32_t tpb = pick_tpb_flat(upb, 128, (uint32_t)std::min<uint64_t>(Q, std::min<uint32_t>(tpb_cap, 384)));
        if (tpb == 0) tpb = pick_tpb_flat(upb, upb, (uint32_t)std::min<uint64_t>(Q, tpb_cap));
        if (tpb) {
            const uint32_t per_sm = (uint32_t)std::max<long>(1, env_long("CSGN_MUL_FLAT_CTAS_PER_SM", 4));
            const uint32_t cap = (uint32_t)std::min<uint64_t>(grid_cap, (uint64_t)dp.sm_count * per_sm);
            cudaError_t err;
            switch (U) {
                case 8: err = launch_flat_u<VT, 8>(fold_mode, a, T1, b, (uint32_t)Q, upb, out, tpb, cap, fold, stream); break;
                case 4: err = launch_flat_u<VT, 4>(fold_mode, a, T1, b, (uint32_t)Q, upb, out, tpb, cap, fold, stream); break;
                case 2: err = launch_flat_u<VT, 2>(fold_mode, a, T1, b, (uint32_t)Q, upb, out, tpb, cap, fold, stream); break;
                default: err = launch_flat_u<VT, 1>(fold_mode, a, T1, b, (uint32_t)Q, upb, out, tpb, cap, fold, stream); break;
            }
            return err;
        }
    }

#endif  // CSGN_BUILD_VARIANTS

    // ---- tiled kernel.  Many small work items balance best, but an item must keep R >= 3 rows per
    // load of its b tile or the L2 re-reads show.  Units per thread: 1 for rows up to 128 KB
    // (chains: many rows of a few hundred blocks), 2 beyond, 4 for products of a GiB and more
    // with long rows.
    const uint64_t Q16 = Q * sizeof(VT) / 16;                     // thresholds were tuned in 16-byte units
    int U = (int)env_long("CSGN_MUL_U", Q16 < 8192 ? 1 : (huge && Q16 >= 32768) ? 4 : 2);
    // lane-aligned fused kernels (blocks of up to 16 units): 4 units per thread cost 79 registers -- three resident CTAs
    // instead of four -- and measured slower than 2 units with more rows per item; long blocks keep 4 (63 registers)
    if (fold_mode && upb <= 16 && env_long("CSGN_MUL_U", 0) <= 0) U = std::min(U, 2);
    if (fold_mode && dbl_words) U = std::min(U, 2);                            // fused double blocks: rows x units <= 32
    // fused, long blocks (32 units and more; the shared-memory fold): wide tiles of few rows -- 4 units per thread and
    // 3 rows per item (B200, tools/r2_sweep.py longfused, profiles/r2_longfused.log: Context(16383,64) 300x300 0.85 -> 0.93
    // of the copy peak, N = 8191 400x400 0.88 -> 0.97; 5 rows and no cap on the grid from 400 MB: 1000x300 0.95 -> 1.05)
    // (unit counts that divide the 256-thread CTA -- N = 4096, 8191, 16383, 32767; for the others the rule before measured
    // as good or better in a batch: 47 units 0.98 / 1.00, 94 units 0.93 / 1.01, 258 units 0.65 / 0.68)
    const bool long_fold = fold_mode && upb >= 32 && 256u % upb == 0 && dbl_words == 0 && Q16 >= 8192 &&
                           env_long("CSGN_MUL_U", 0) <= 0 && env_long("CSGN_MUL_LONGFOLD", 1) != 0;
    const bool mid = out_units * sizeof(VT) >= (400ull << 20);    // 400 MB of output and more
    if (long_fold) U = 4;
    U = U >= 8 ? 8 : U >= 4 ? 4 : U >= 2 ? 2 : 1;
    while (U > 1 && (uint64_t)upb * U > Q) U >>= 1;
    const uint32_t r_smem = std::max<uint32_t>(1, std::min<uint32_t>(64, kMulMaxSmem / (upb * (uint32_t)sizeof(VT))));
    const uint64_t target_items =
        (uint64_t)dp.sm_count * (uint64_t)env_long("CSGN_MUL_ITEMS_PER_SM", huge ? 128 : 32);
    // Fused kernels for blocks of up to 16 units run lane-aligned (whole blocks per warp: the fold is two redux.sync
    // per item, no shared memory, no barrier); their CTA is a whole number of warps -- unless the row is so short that
    // whole warps would pad the column tiles by more than a few percent (chains: 250 or 1250 units per row).
    bool align = fold_mode != 0 && upb <= 16 && dbl_words == 0 && env_long("CSGN_MUL_ALIGN", 1) != 0;
    const uint32_t lanes_used = align ? (32u / upb) * upb : 32u;
    uint32_t tpb = 0;
    uint64_t R = 1;
    uint64_t step_units = 0;            // units one row step of the CTA covers
    for (;; U >>= 1) {
        uint32_t tpb_free = pick_tpb(upb, Q, U, tpb_cap);
        if (tpb_free == 0) tpb_free = upb;
        tpb = tpb_free;
        step_units = tpb_free;
        if (align) {
            // the warp count (4 .. cap/32, nearest to 8 on ties) that pads the last column tile least
            double best = 1e30;
            uint32_t best_w = 0;
            const uint32_t max_warps = std::max<uint32_t>(1, tpb_cap / 32);
            for (uint32_t w = std::min<uint32_t>(4, max_warps); w <= max_warps; ++w) {
                const uint64_t tile = (uint64_t)w * lanes_used * U, nt = (Q + tile - 1) / tile;
                const double pad = (double)(nt * tile) / (double)Q * (32.0 / lanes_used);
                const bool closer = best_w == 0 || (w > 8 ? w - 8 : 8 - w) < (best_w > 8 ? best_w - 8 : 8 - best_w);
                if (pad < best - 1e-9 || (pad < best + 1e-9 && closer)) {
                    best = pad;
                    best_w = w;
                }
            }
            const uint64_t tile_f = (uint64_t)tpb_free * U, nt_f = (Q + tile_f - 1) / tile_f;
            const double pad_free = (double)(nt_f * tile_f) / (double)Q * ((double)((tpb_free + 31) / 32 * 32) / tpb_free);
            if (env_long("CSGN_MUL_ALIGN", 1) < 2 && best > 1.06 * pad_free) align = false;      // 2: lane-aligned whatever it pads
            else {
                tpb = best_w * 32;
                step_units = (uint64_t)best_w * lanes_used;
            }
        }
        const uint64_t n_col_tiles = (Q + step_units * U - 1) / (step_units * U);
        R = (T1 * n_col_tiles + target_items - 1) / target_items;
        if (R >= 3 || U == 1 || fold_mode || env_long("CSGN_MUL_U", 0) > 0) break;
    }
    // B200 sweeps (tools/r2_sweep.py rsel, profiles/r2_rsel*.log).  Chains -- very many short rows -- write fastest when
    // the resident CTAs cover a compact window of the product: few rows per item.  A fused item carries fixed work (its
    // slice of b, the staged rows, two barriers, the fold) that only pays off over ~56 KB of product: 8 rows of a
    // 7.7 KB column tile at N=1247, two dozen rows where a tile row is short (chains, N=191).
    if (huge && U == 1 && !fold_mode) R = std::min<uint64_t>(R, 4);
    if (fold_mode) {
        const uint64_t row_bytes = step_units * (uint64_t)U * sizeof(VT);
        R = std::max<uint64_t>(6, (56 * 1024 + row_bytes - 1) / row_bytes);
        if (long_fold && U == 4 && !huge) R = mid ? 5 : 3;
    }
    R = (uint64_t)env_long("CSGN_MUL_R", (long)R);
    uint32_t r_max = r_smem;
    if (fold_mode) r_max = std::min<uint32_t>(r_max, (dbl_words ? 32u : 64u) / (uint32_t)U);   // one 64-bit fail word per thread and item
    R = std::max<uint64_t>(1, std::min<uint64_t>(R, std::min<uint64_t>(r_max, T1)));

    // A fused CTA ends with a global atomic and a ticket (fold.cuh): with one item per CTA every item pays that round
    // trip.  The shared-memory-fold kernels (long blocks) therefore run as a persistent grid of 8 CTAs per SM that loop
    // over the items (B200, tools/fused_grid_sweep.py: Context(16383,64) 300x300 fused 31.7 -> 28.6 us in a batch,
    // 35.6 -> 33.3 us alone); the lane-aligned kernels (N = 1247) measured best with one CTA per item in a batch.
    uint64_t grid_cap_f = grid_cap;
    if (fold_mode && !align && dbl_words == 0 && upb >= 32 && !huge && !(long_fold && mid))    // (a GiB and more: 1.11 -> 0.98 with the cap)
        grid_cap_f = std::min<uint64_t>(grid_cap, (uint64_t)dp.sm_count * (uint64_t)env_long("CSGN_MUL_FOLD_CTAS_PER_SM", 8));
    // ... and 16 CTAs per SM when it runs alone (no other kernel fills the SMs behind its tail): 27.2 -> 25.4 us
    if (fold_mode && align && !huge && fold && !fold->overlapped)
        grid_cap_f = std::min<uint64_t>(grid_cap, (uint64_t)dp.sm_count * (uint64_t)env_long("CSGN_MUL_FOLD_CTAS_PER_SM_ALONE", 16));
    switch (U) {
        case 8: return launch_tiled_u<VT, 8>(fold_mode, align, a, T1, b, Q, upb, out, tpb, (uint32_t)R, grid_cap_f, fold, stream, dbl_words);
        case 4: return launch_tiled_u<VT, 4>(fold_mode, align, a, T1, b, Q, upb, out, tpb, (uint32_t)R, grid_cap_f, fold, stream, dbl_words);
        case 2: return launch_tiled_u<VT, 2>(fold_mode, align, a, T1, b, Q, upb, out, tpb, (uint32_t)R, grid_cap_f, fold, stream, dbl_words);
        default: return launch_tiled_u<VT, 1>(fold_mode, align, a, T1, b, Q, upb, out, tpb, (uint32_t)R, grid_cap_f, fold, stream, dbl_words);
    }
}

}  // namespace

bool mul_fold_supported(uint32_t L) { return L != 0 && ((L & 1u) ? L : L / 2) <= (uint32_t)kMulMaxThreads; }

cudaError_t launch_mul(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2, uint32_t L,
                       uint64_t *out, cudaStream_t stream, const MulFold *fold) {
    if (T1 == 0 || T2 == 0 || L == 0) return cudaSuccess;
    const int fold_mode = fold ? (out ? 1 : 2) : 0;
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                             reinterpret_cast<uintptr_t>(out) | (fold ? reinterpret_cast<uintptr_t>(fold->mask) : 0)) & 15u) == 0;
    const bool units16 = !(L & 1u) && aligned16;
    const uint32_t upb = units16 ? L / 2 : L;
    if (upb > (uint32_t)kMulMaxThreads || env_long("CSGN_MUL_GENERIC", 0)) {
        if (fold) return cudaErrorNotSupported;      // the caller multiplies, then folds (two launches)
        const DeviceProps &dp = device_props();
        const uint64_t row_words = T2 * L, total = T1 * row_words;
        const uint64_t want = (total + 255) / 256;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)dp.sm_count * 16);
        count_launch();
        return launch_kernel(mul_outer_generic_kernel, grid, 256, 0, stream, a, b, out, L, row_words, total);
    }
    if (T2 == 1 && T1 > 1 && !env_long("CSGN_MUL_NOSWAP", 0)) {
        // a (T1 blocks) x one block is, as a word stream, that block x a (out[i] = a_i & b_0
        // either way) -- and the swapped form is one long row instead of T1 tiny ones
        return launch_mul(b, 1, a, T1, L, out, stream, fold);
    }
    // Odd L (half of all contexts): a block is not a whole number of 16-byte units, but TWO blocks are.  With an even
    // number of right-operand blocks the plain multiply reads b as T2/2 double blocks of 2L words and stages every row of
    // a as a_i || a_i -- the same words out, with 16-byte loads and stores instead of 8-byte ones.
    // The fused kernel does the same, with two verdicts per double block.
    const bool dbl = (L & 1u) && (T2 % 2 == 0) && L <= (uint32_t)kMulMaxThreads && env_long("CSGN_MUL_DOUBLE", 1) &&
                     ((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0 &&
                     (!fold || (L > 16 && (reinterpret_cast<uintptr_t>(fold->mask) & 7u) == 0));   // short odd blocks fuse lane-aligned on 8-byte units
    const cudaError_t err = dbl       ? launch_units<uint4>(fold_mode, a, T1, b, T2 / 2, L, out, fold, stream, L)
                            : units16 ? launch_units<uint4>(fold_mode, a, T1, b, T2, upb, out, fold, stream)
                                      : launch_units<uint2>(fold_mode, a, T1, b, T2, upb, out, fold, stream);
    count_launch();
    return err;
}

}  // namespace csgn
