// mul.cu -- K1, ciphertext multiply: the all-pairs AND of T1 x T2 blocks.
//
//   out[(i*T2+j)*L+k] = a[i*L+k] & b[j*L+k]     (reference src/Ciphertext.cpp:153-163;
//                                                the 1x1 shortcut :124-131 is T1=T2=1)
//
// Roofline: HBM WRITE bandwidth.  8*L bytes are written per output block, the
// operands (8*L*(T1+T2) bytes in total) stay L2-resident, one 64-bit AND per 8 bytes.
//
// Shape of the kernel.  Output row i is the whole right operand, seen as one flat
// stream of Q = T2*L/2 16-byte units, ANDed with block a_i repeated with period
// L4 = L/2 units.  The CTA size is a multiple of L4, so a thread that walks the flat
// stream with stride blockDim always meets the same 16-byte fragment of a_i:
//   - a work item is (column tile of U*blockDim units) x (chunk of R rows);
//   - the thread's U units of b are loaded ONCE into registers (coalesced);
//   - the R blocks of a are staged in shared memory (R*L4*16 bytes);
//   - per row: one LDS.128 for the thread's fragment of a_i, U ANDs, U coalesced
//     128-bit streaming stores (st.global.cs -- nothing re-reads the product).
// b is re-read from L2 once per R rows, so L2 read traffic is 1/R of the write
// stream; no integer division happens inside the row loop.
#include "kernels.cuh"
#include "launch.cuh"

#include <algorithm>

namespace csgn {
namespace {

constexpr int kMulMaxThreads = 512;
constexpr uint32_t kMulMaxSmem = 32 * 1024;

__device__ __forceinline__ uint4 and4(const uint4 a, const uint4 b) {
    return make_uint4(a.x & b.x, a.y & b.y, a.z & b.z, a.w & b.w);
}

// Store flavours (tuning): 0 = st.global.cs (evict-first), 1 = default write-back,
// 2 = st.global.wt (write-through).
template <int MODE>
__device__ __forceinline__ void st_out(uint4 *p, const uint4 v) {
    if (MODE == 0) __stcs(p, v);
    else if (MODE == 1) *p = v;
    else __stwt(p, v);
}

template <int U, int MODE>
__global__ void __launch_bounds__(kMulMaxThreads)
mul_outer_kernel(const uint4 *__restrict__ A4, const uint4 *__restrict__ B4, uint4 *__restrict__ out4,
                 const uint32_t L4, const uint64_t T1, const uint64_t Q, const uint32_t R,
                 const uint32_t n_col_tiles, const uint64_t n_items, const uint32_t pf_chunks) {
    extern __shared__ uint4 sA[];
    const uint32_t tpb = blockDim.x;
    const uint32_t k4 = threadIdx.x % L4;
    const uint64_t tile_q = (uint64_t)tpb * U;
    pdl_enter();

    // In real use the left operand is DRAM-cold (the previous product flushed L2) and an
    // item is short, so a cold a-chunk is a full DRAM latency in front of every item.  CTAs
    // start in blockIdx order, so CTA j warms L2 for the items that start ~two waves later:
    // the chunk of item j + pf_chunks*n_col_tiles (done by the column-0 CTA of each chunk);
    // the first pf_chunks chunks, which nobody is ahead of, are requested line by line by
    // the first CTAs.
    if (pf_chunks) {
        const uint64_t a_bytes = T1 * L4 * 16u;
        const uint64_t head_bytes = min(a_bytes, (uint64_t)pf_chunks * R * L4 * 16u);
        if (threadIdx.x == 0 && (uint64_t)blockIdx.x * 128u < head_bytes)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A4) + (uint64_t)blockIdx.x * 128u));
        if (blockIdx.x % n_col_tiles == 0) {
            const uint64_t row0 = ((uint64_t)blockIdx.x / n_col_tiles + pf_chunks) * R;
            if (row0 < T1) {
                const uint32_t bytes = (uint32_t)min((uint64_t)R, T1 - row0) * L4 * 16u;
                const uint32_t off = threadIdx.x * 128u;
                if (off < bytes + 128u)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(A4 + row0 * L4) +
                                                                   min(off, bytes - 1u)));
            }
        }
    }

    for (uint64_t item = blockIdx.x; item < n_items; item += gridDim.x) {
        const uint32_t ct = (uint32_t)(item % n_col_tiles);
        const uint64_t row0 = (item / n_col_tiles) * R;
        const uint32_t nrows = (uint32_t)min((uint64_t)R, T1 - row0);
        const uint64_t q0 = (uint64_t)ct * tile_q + threadIdx.x;

        uint4 b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t q = q0 + (uint64_t)u * tpb;
            b[u] = (q < Q) ? __ldg(B4 + q) : make_uint4(0u, 0u, 0u, 0u);
        }

        __syncthreads();  // the previous item's readers of sA are done
        const uint4 *a_chunk = A4 + row0 * L4;
        for (uint32_t idx = threadIdx.x; idx < nrows * L4; idx += tpb) sA[idx] = __ldg(a_chunk + idx);
        __syncthreads();

        uint4 *o = out4 + row0 * Q + q0;
        const uint4 *sa = sA + k4;
        if ((uint64_t)(ct + 1) * tile_q <= Q) {
#pragma unroll 4
            for (uint32_t r = 0; r < nrows; ++r, o += Q, sa += L4) {
                const uint4 a = *sa;
#pragma unroll
                for (int u = 0; u < U; ++u) st_out<MODE>(o + (uint64_t)u * tpb, and4(a, b[u]));
            }
        } else {
            // ragged last column tile: per-unit bounds, loop-invariant predicates
            bool live[U];
#pragma unroll
            for (int u = 0; u < U; ++u) live[u] = q0 + (uint64_t)u * tpb < Q;
            for (uint32_t r = 0; r < nrows; ++r, o += Q, sa += L4) {
                const uint4 a = *sa;
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (live[u]) st_out<MODE>(o + (uint64_t)u * tpb, and4(a, b[u]));
            }
        }
    }
}


// Any L (odd included), any alignment: one 64-bit word per thread-iteration.
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

template <int U, int MODE>
cudaError_t launch_v4m(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2, uint32_t L4,
                       uint64_t *out, uint32_t tpb, uint32_t R, uint64_t grid_cap, cudaStream_t stream) {
    const uint64_t Q = T2 * L4;
    const uint64_t tile_q = (uint64_t)tpb * U;
    const uint64_t n_col_tiles = (Q + tile_q - 1) / tile_q;
    const uint64_t n_chunks = (T1 + R - 1) / R;
    const uint64_t n_items = n_col_tiles * n_chunks;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(n_items, grid_cap);
    const size_t smem = (size_t)R * L4 * sizeof(uint4);
    // prefetch distance: the chunks that ~two waves of resident CTAs cover
    const uint64_t ahead_items = (uint64_t)device_props().sm_count * (uint64_t)env_long("CSGN_MUL_PF_CTAS_PER_SM", 8);
    const uint32_t pf_chunks =
        env_long("CSGN_MUL_PF_CTAS_PER_SM", 8) > 0 ? (uint32_t)((ahead_items + n_col_tiles - 1) / n_col_tiles) : 0u;
    return launch_kernel(mul_outer_kernel<U, MODE>, grid, tpb, smem, stream, reinterpret_cast<const uint4 *>(a),
                         reinterpret_cast<const uint4 *>(b), reinterpret_cast<uint4 *>(out), L4, T1, Q, R,
                         (uint32_t)n_col_tiles, n_items, pf_chunks);
}

template <int U>
cudaError_t launch_v4(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2, uint32_t L4,
                      uint64_t *out, uint32_t tpb, uint32_t R, uint64_t grid_cap, cudaStream_t stream) {
    switch (env_long("CSGN_MUL_STORE", 0)) {
        case 1: return launch_v4m<U, 1>(a, T1, b, T2, L4, out, tpb, R, grid_cap, stream);
        case 2: return launch_v4m<U, 2>(a, T1, b, T2, L4, out, tpb, R, grid_cap, stream);
        default: return launch_v4m<U, 0>(a, T1, b, T2, L4, out, tpb, R, grid_cap, stream);
    }
}

// CTA size for the tiled kernel: a multiple of L4 in [192, cap].  A partial last warp costs
// issue slots in every item; a ragged last column tile only makes that tile's items shorter
// (items are scheduled dynamically), so it weighs a quarter.
uint32_t pick_tpb(uint32_t L4, uint64_t Q, int U, uint32_t cap) {
    uint32_t best = 0;
    double best_cost = 1e30;
    for (uint32_t t = (cap / L4) * L4; t >= L4 && t > 0; t -= L4) {
        const uint64_t tile = (uint64_t)t * U;
        const uint64_t n_tiles = (Q + tile - 1) / tile;
        const double pad = (double)(n_tiles * tile) / (double)Q;
        const double warp = (double)((t + 31) / 32 * 32) / (double)t;
        const double cost = warp * (1.0 + 0.25 * (pad - 1.0)) * (t < 256 ? 1.03 : 1.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = t;
        }
        if (t < 192 + L4) break;
    }
    return best;
}

}  // namespace

cudaError_t launch_mul(const uint64_t *a, uint64_t T1, const uint64_t *b, uint64_t T2, uint32_t L,
                       uint64_t *out, cudaStream_t stream) {
    if (T1 == 0 || T2 == 0 || L == 0) return cudaSuccess;
    const DeviceProps &dp = device_props();
    const bool aligned = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                           reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
    const uint32_t L4 = L / 2;
    if ((L & 1u) || !aligned || L4 > (uint32_t)kMulMaxThreads || env_long("CSGN_MUL_GENERIC", 0)) {
        const uint64_t row_words = T2 * L, total = T1 * row_words;
        const uint64_t want = (total + 255) / 256;
        const uint32_t grid = (uint32_t)std::min<uint64_t>(want, (uint64_t)dp.sm_count * 16);
        count_launch();
        return launch_kernel(mul_outer_generic_kernel, grid, 256, 0, stream, a, b, out, L, row_words, total);
    }

    const uint64_t Q = T2 * L4;
    if (T2 == 1 && T1 > 1 && !env_long("CSGN_MUL_NOSWAP", 0)) {
        // a (T1 blocks) x one block is, as a word stream, that block x a (out[i] = a_i & b_0
        // either way) -- and the swapped form is one long row instead of T1 tiny ones
        return launch_mul(b, 1, a, T1, L, out, stream);
    }
    const uint32_t tpb_cap = (uint32_t)env_long("CSGN_MUL_TPB", 512);
    // Tiled kernel.  Many small work items balance best, but an item must keep R >= 3 rows per
    // load of its b tile or the L2 re-reads show.  Units per thread: 1 for rows up to 128 KB
    // (chains: many rows of a few hundred blocks), 2 beyond, 4 for products of a GiB and more
    // with long rows.
    const bool huge = T1 * Q >= (1ull << 26);     // >= 1 GiB of output
    int U = (int)env_long("CSGN_MUL_U", Q < 8192 ? 1 : (huge && Q >= 32768) ? 4 : 2);
    U = U >= 8 ? 8 : U >= 4 ? 4 : U >= 2 ? 2 : 1;
    while (U > 1 && (uint64_t)L4 * U > Q) U >>= 1;
    const uint32_t r_max = std::max<uint32_t>(1, std::min<uint32_t>(64, kMulMaxSmem / (L4 * 16)));
    const uint64_t target_items =
        (uint64_t)dp.sm_count * (uint64_t)env_long("CSGN_MUL_ITEMS_PER_SM", huge ? 128 : 32);
    uint32_t tpb = 0;
    uint64_t R = 1;
    for (;; U >>= 1) {
        tpb = pick_tpb(L4, Q, U, tpb_cap);
        if (tpb == 0) tpb = L4;
        const uint64_t n_col_tiles = (Q + (uint64_t)tpb * U - 1) / ((uint64_t)tpb * U);
        R = (T1 * n_col_tiles + target_items - 1) / target_items;
        if (R >= 3 || U == 1 || env_long("CSGN_MUL_U", 0) > 0) break;
    }
    R = (uint64_t)env_long("CSGN_MUL_R", (long)R);
    R = std::max<uint64_t>(1, std::min<uint64_t>(R, std::min<uint64_t>(r_max, T1)));
    const uint64_t grid_cap = (uint64_t)env_long("CSGN_MUL_GRID", 1 << 30);

    cudaError_t err;
    switch (U) {
        case 8: err = launch_v4<8>(a, T1, b, T2, L4, out, tpb, (uint32_t)R, grid_cap, stream); break;
        case 4: err = launch_v4<4>(a, T1, b, T2, L4, out, tpb, (uint32_t)R, grid_cap, stream); break;
        case 2: err = launch_v4<2>(a, T1, b, T2, L4, out, tpb, (uint32_t)R, grid_cap, stream); break;
        default: err = launch_v4<1>(a, T1, b, T2, L4, out, tpb, (uint32_t)R, grid_cap, stream); break;
    }
    count_launch();
    return err;
}

}  // namespace csgn
