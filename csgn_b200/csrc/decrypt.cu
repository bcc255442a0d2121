// decrypt.cu -- K3, the decrypt fold.
//
//   count = #{ blocks k : for every word w, (v[k*L+w] & M[w]) == M[w] }
//   decrypt = count & 1                       (reference src/SecretKey.cpp:126-140)
//
// M is the secret key as a position mask (bit 63-(s&63) of word s>>6 per secret
// position s).  The reference first unpacks every bit to one byte
// (src/SecretKey.cpp:113-124) and then ANDs the D selected bytes; testing the packed
// words against the mask is the same predicate.
//
// Roofline: HBM READ bandwidth, 8*L bytes per block, one bit out.
//
// The ciphertext is read as one flat stream of units -- 16 bytes (uint4) when L is even and the
// words are 16-byte aligned, 8 bytes (uint2) otherwise (odd L: N = 191, 4097 ...) -- fully coalesced.
// Shapes of the same fold, chosen by units per block (UPB):
//   lanes  UPB <= 16          a warp step covers 32/UPB whole blocks; a lane always holds the same unit of
//                             a block, its mask unit lives in registers, a block's verdict is UPB adjacent
//                             bits of one ballot                                  (N=1247: UPB = 10)
//   wide   UPB % 32 == 0      warp per block, UPB/32 coalesced loads per lane, one vote per block
//                                                                                 (N=16383: UPB = 128)
//   string any UPB <= 512     a warp owns 32 consecutive blocks and walks them in UPB coalesced steps; each
//                             step's ballot is 32 bits of a per-chunk fail string in shared memory; block b
//                             is satisfied iff bits [b*UPB, (b+1)*UPB) are clear -- lane b checks exactly that
//   window odd L (3, 17..999 words) and even L with 17..128 units that fit neither of the first two: the block
//                             structure is taken off the load pattern -- warps stream contiguous runs with every
//                             lane loading 16 bytes, and the blocks are laid over each step's two ballots afterwards
//   rows   long blocks whose unit count is not a multiple of 32                   (N=33000: UPB = 258)
// plus a warp-per-block kernel for blocks longer than 512 units.  Counts fold lane -> warp -> CTA -> ONE
// atomic per CTA (fold.cuh); the last CTA publishes the total, so a decrypt is ONE launch.
#include "kernels.cuh"
#include "launch.cuh"
#include "bulk.cuh"
#include "peer.cuh"
#include "fold.cuh"

#include <algorithm>
#include <cstring>

namespace csgn {
namespace {

constexpr int kDecThreads = 256;
constexpr int kDecWarps = kDecThreads / 32;
constexpr uint32_t kDecMaxUnits = 512;   // mask x2 + fail strings stay within 32 KB of shared memory

template <typename VT> __device__ __forceinline__ VT vzero();
template <> __device__ __forceinline__ uint4 vzero<uint4>() { return make_uint4(0u, 0u, 0u, 0u); }
template <> __device__ __forceinline__ uint2 vzero<uint2>() { return make_uint2(0u, 0u); }
template <typename VT> __device__ __forceinline__ VT ld_stream(const VT *p) { return __ldcs(p); }

// ---------------------------------------------------------------------------------------
// string: any units-per-block up to kDecMaxUnits.  UPBC > 0: known at compile time; 0: runtime.
// ---------------------------------------------------------------------------------------
template <typename VT, int UPBC, int UNROLL, int MINB>
__global__ void __launch_bounds__(kDecThreads, MINB)
decrypt_count_kernel(const VT *__restrict__ V, const uint64_t T, const uint32_t upb_rt, const VT *__restrict__ M,
                     const uint32_t cpw, uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t UPB = UPBC ? (uint32_t)UPBC : upb_rt;
    VT *sM2 = reinterpret_cast<VT *>(smem_raw);                      // mask, twice over
    uint32_t *sF = reinterpret_cast<uint32_t *>(sM2 + 2 * UPB);      // fail strings
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    // The key mask is not produced by the previous kernel of the stream: it is staged BEFORE the PDL wait, so this
    // prologue (a DRAM miss after a multiply has streamed through L2) overlaps the predecessor's tail.
    for (uint32_t i = threadIdx.x; i < 2 * UPB; i += blockDim.x) sM2[i] = M[i < UPB ? i : i - UPB];
    __syncthreads();
    pdl_enter();

    uint32_t *sFw = sF + warp * UPB;
    const VT *mk = sM2 + (lane % UPB);
    const uint32_t step = 32u % UPB;          // advance of the unit-in-block index per walk step
    const uint64_t n_units = T * UPB;
    const uint64_t n_chunks = (T + 31) >> 5;
    uint64_t my_count = 0;

    // cpw == 0: persistent warps sweep the stream together, round k covers chunks [k*W, (k+1)*W).
    // cpw  > 0: many short CTAs, each owning cpw*8 consecutive chunks, placed by the hardware
    //           scheduler as SMs free up (SMs do not all stream at the same speed).
    const uint64_t first = cpw ? (uint64_t)blockIdx.x * kDecWarps * cpw + warp : (uint64_t)blockIdx.x * kDecWarps + warp;
    const uint64_t stride = cpw ? (uint64_t)kDecWarps : (uint64_t)gridDim.x * kDecWarps;
    const uint64_t last = cpw ? min(n_chunks, ((uint64_t)blockIdx.x + 1) * kDecWarps * cpw) : n_chunks;
    for (uint64_t chunk = first; chunk < last; chunk += stride) {
        const VT *src = V + (chunk * 32u * UPB + lane);
        const bool full = (chunk + 1) * 32u <= T;    // warp-uniform: no per-load bounds in the common case
        uint32_t koff = 0;                           // (32*r) % UPB
        for (uint32_t r = 0; r < UPB; r += UNROLL) {
            VT v[UNROLL];
            if (full) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u)
                    if ((UPBC && UPBC % UNROLL == 0) || r + u < UPB) v[u] = ld_stream(src + 32u * (r + u));
            } else {
                const uint64_t q_lane = chunk * 32u * UPB + lane;
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) {
                    const uint64_t q = q_lane + 32u * (r + u);
                    v[u] = (r + u < UPB && q < n_units) ? ld_stream(V + q) : vzero<VT>();
                }
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if ((UPBC && UPBC % UNROLL == 0) || r + u < UPB) {   // warp-uniform
                    const bool f = unit_fails(v[u], mk[koff]);
                    const uint32_t bal = __ballot_sync(0xffffffffu, f);
                    if (lane == 0) sFw[r + u] = bal;
                    koff += step;
                    if (koff >= UPB) koff -= UPB;
                }
            }
        }
        __syncwarp();
        // lane b <-> block b of the chunk: bits [b*UPB, b*UPB+UPB) of the fail string
        const uint64_t blk = chunk * 32u + lane;
        const uint32_t lo = lane * UPB, hi = lo + UPB;
        uint32_t any = 0;
        for (uint32_t w = lo >> 5; w <= (hi - 1) >> 5; ++w) {
            const uint32_t first = max(lo, w << 5) - (w << 5);
            const uint32_t last = min(hi, (w + 1) << 5) - (w << 5);   // exclusive, 1..32
            const uint32_t m = (last - first == 32u) ? 0xffffffffu : (((1u << (last - first)) - 1u) << first);
            any |= sFw[w] & m;
        }
        my_count += (blk < T && any == 0u) ? 1u : 0u;
        __syncwarp();                        // before the next chunk overwrites sFw
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}


// ---------------------------------------------------------------------------------------
// string over DOUBLE blocks, for odd L of 17 words and more: two consecutive blocks are L 16-byte units, so the walk
// above runs on 16-byte loads over the T/2 double blocks; a unit's low word belongs to the first block of its pair when
// its word index 2k is below L, its high word when 2k+1 is -- two ballots per step, two fail strings per warp, two
// verdicts per lane.  An odd last block is folded by the grid's first warp with 8-byte loads.
// ---------------------------------------------------------------------------------------
template <int UNROLL, int MINB>
__global__ void __launch_bounds__(kDecThreads, MINB)
decrypt_count_string_pairs_kernel(const uint4 *__restrict__ V, const uint64_t T, const uint32_t L,
                                  const uint64_t *__restrict__ M, uint64_t *scratch, uint64_t *count_out,
                                  const __grid_constant__ PeerPush pp) {
    extern __shared__ uint4 smem_raw[];
    const uint32_t UPB = L;                                          // units per double block
    uint4 *sM2 = smem_raw;                                           // the double block's mask units, twice over
    uint32_t *sF = reinterpret_cast<uint32_t *>(sM2 + 2 * UPB);      // fail strings: first blocks, second blocks
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 2 * UPB; i += blockDim.x) {
        const uint32_t u = i < UPB ? i : i - UPB;
        const uint64_t lo = M[2u * u < L ? 2u * u : 2u * u - L], hi = M[2u * u + 1u < L ? 2u * u + 1u : 2u * u + 1u - L];
        sM2[i] = make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
    }
    __syncthreads();
    pdl_enter();

    uint32_t *sF0 = sF + warp * 2u * UPB, *sF1 = sF0 + UPB;
    const uint32_t k0 = lane % UPB;
    const uint32_t step = 32u % UPB;
    const uint64_t n_pairs = T >> 1;
    const uint64_t n_units = n_pairs * UPB;
    const uint64_t n_chunks = (n_pairs + 31) >> 5;
    uint64_t my_count = 0;
    for (uint64_t chunk = (uint64_t)blockIdx.x * kDecWarps + warp; chunk < n_chunks; chunk += (uint64_t)gridDim.x * kDecWarps) {
        const uint64_t q_lane = chunk * 32u * UPB + lane;
        const bool full = (chunk + 1) * 32u <= n_pairs;
        uint32_t koff = 0;
        for (uint32_t r = 0; r < UPB; r += UNROLL) {
            uint4 v[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                const uint64_t q = q_lane + 32u * (r + u);
                v[u] = (r + u < UPB && (full || q < n_units)) ? ld_stream(V + q) : vzero<uint4>();
            }
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (r + u < UPB) {                                   // warp-uniform
                    uint32_t k = k0 + koff;
                    if (k >= UPB) k -= UPB;
                    const uint4 m = sM2[k];
                    const bool fl = ((~v[u].x & m.x) | (~v[u].y & m.y)) != 0u, fh = ((~v[u].z & m.z) | (~v[u].w & m.w)) != 0u;
                    const bool lo_first = 2u * k < L, hi_first = 2u * k + 1u < L;
                    const uint32_t b0 = __ballot_sync(0xffffffffu, (lo_first && fl) || (hi_first && fh));
                    const uint32_t b1 = __ballot_sync(0xffffffffu, (!lo_first && fl) || (!hi_first && fh));
                    if (lane == 0) {
                        sF0[r + u] = b0;
                        sF1[r + u] = b1;
                    }
                    koff += step;
                    if (koff >= UPB) koff -= UPB;
                }
            }
        }
        __syncwarp();
        // lane b <-> double block b of the chunk: bits [b*UPB, b*UPB+UPB) of each fail string
        const uint64_t pair = chunk * 32u + lane;
        const uint32_t lo = lane * UPB, hi = lo + UPB;
        uint32_t any0 = 0, any1 = 0;
        for (uint32_t w = lo >> 5; w <= (hi - 1) >> 5; ++w) {
            const uint32_t first = max(lo, w << 5) - (w << 5);
            const uint32_t last = min(hi, (w + 1) << 5) - (w << 5);   // exclusive, 1..32
            const uint32_t m = (last - first == 32u) ? 0xffffffffu : (((1u << (last - first)) - 1u) << first);
            any0 |= sF0[w] & m;
            any1 |= sF1[w] & m;
        }
        if (pair < n_pairs) my_count += (any0 == 0u ? 1u : 0u) + (any1 == 0u ? 1u : 0u);
        __syncwarp();                        // before the next chunk overwrites the strings
    }
    if ((T & 1ull) && blockIdx.x == 0 && warp == 0) {
        const uint64_t *last = reinterpret_cast<const uint64_t *>(V) + (T - 1) * L;
        bool f = false;
        for (uint32_t w = lane; w < L; w += 32) f |= (~__ldcs(last + w) & __ldg(M + w)) != 0ull;
        const bool bad = __any_sync(0xffffffffu, f);
        if (lane == 0 && !bad) ++my_count;
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// ---------------------------------------------------------------------------------------
// lanes: short blocks (UPB <= 16 units, e.g. N=1247: 10 units of 16 bytes).
//
// A warp step covers BPS = 32 / UPB whole blocks with the first BPS*UPB lanes (30 of 32 at UPB = 10;
// the step is still one contiguous, sector-aligned run of 480 bytes).  A lane therefore always
// sits on the SAME unit of a block: its key-mask unit lives in registers for the whole launch (taken
// from the kernel parameters -- no global or shared load at all for the mask), a block's verdict is
// UPB adjacent bits of one ballot, and every lane computes the same count from it -- no fail string in
// shared memory, no __syncwarp, no per-unit mask lookup.  What is left per step is one load, a few
// logic ops, a vote and a handful of uniform integer ops, so U independent loads per lane stay in
// flight with registers to spare.
// ---------------------------------------------------------------------------------------
template <typename VT, int UPB, int U, int MINB>
__global__ void __launch_bounds__(kDecThreads, MINB)
decrypt_count_lanes_kernel(const VT *__restrict__ V, const uint64_t T, const __grid_constant__ ParamMask pmask,
                           uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    constexpr uint32_t BPS = 32u / UPB, ACTIVE = BPS * UPB;
    constexpr uint32_t BLKMASK = UPB == 32 ? 0xffffffffu : (1u << UPB) - 1u;
    const uint32_t lane = threadIdx.x & 31u;
    const bool active = lane < ACTIVE;
    const VT m = active ? reinterpret_cast<const VT *>(&pmask)[lane % UPB] : vzero<VT>();   // parameters: nothing to wait for
    pdl_enter();

    const uint64_t n_steps = (T + BPS - 1) / BPS;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint32_t cnt = 0;                         // warp-uniform: satisfied blocks seen by this warp
    for (uint64_t base = warp_global * U; base < n_steps; base += n_warps * U) {
        const VT *src = V + (base * ACTIVE + lane);
        VT v[U];
        if ((base + U) * BPS <= T) {          // warp-uniform: all U steps lie inside the ciphertext
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = active ? ld_stream(src + (uint32_t)u * ACTIVE) : vzero<VT>();
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t bal = __ballot_sync(0xffffffffu, unit_fails(v[u], m));
#pragma unroll
                for (uint32_t b = 0; b < BPS; ++b) cnt += ((bal >> (b * UPB)) & BLKMASK) == 0u ? 1u : 0u;
            }
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint64_t blk = (base + u) * BPS + lane / UPB;        // this lane's block
                v[u] = (active && blk < T) ? ld_stream(src + (uint32_t)u * ACTIVE) : vzero<VT>();
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t bal = __ballot_sync(0xffffffffu, unit_fails(v[u], m));
#pragma unroll
                for (uint32_t b = 0; b < BPS; ++b)
                    cnt += ((base + u) * BPS + b < T && ((bal >> (b * UPB)) & BLKMASK) == 0u) ? 1u : 0u;
            }
        }
    }
    fold_and_publish(lane == 0 ? (uint64_t)cnt : 0ull, scratch, count_out, pp);
}

// ---------------------------------------------------------------------------------------
// wide: UPB a multiple of 32 (e.g. N=16383: 128 units): a block is UPL = UPB/32 coalesced warp
// loads; one warp folds BPI blocks per iteration (BPI*UPL independent 16-byte loads in
// flight per lane) and votes once per block.
// ---------------------------------------------------------------------------------------
template <int UPL, int BPI>
__global__ void __launch_bounds__(kDecThreads, 4)
decrypt_count_wide_kernel(const uint4 *__restrict__ V4, const uint64_t T, const uint4 *__restrict__ M4,
                          uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    const uint32_t lane = threadIdx.x & 31u;
    uint4 m[UPL];
#pragma unroll
    for (int u = 0; u < UPL; ++u) m[u] = __ldg(M4 + 32 * u + lane);     // the key is not the predecessor's output
    pdl_enter();
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_groups = (T + BPI - 1) / BPI;
    uint64_t my_count = 0;
    for (uint64_t grp = warp_global; grp < n_groups; grp += n_warps) {
        uint4 v[BPI][UPL];
#pragma unroll
        for (int b = 0; b < BPI; ++b) {
            const uint64_t blk = grp * BPI + b;
            const uint4 *row = V4 + blk * (32u * UPL) + lane;
#pragma unroll
            for (int u = 0; u < UPL; ++u) v[b][u] = (blk < T) ? ld_stream(row + 32 * u) : vzero<uint4>();
        }
#pragma unroll
        for (int b = 0; b < BPI; ++b) {
            bool f = false;
#pragma unroll
            for (int u = 0; u < UPL; ++u) f |= unit_fails(v[b][u], m[u]);
            const bool bad = __any_sync(0xffffffffu, f);
            if (lane == 0 && !bad && grp * BPI + b < T) ++my_count;
        }
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// ---------------------------------------------------------------------------------------
// Long blocks of ANY length (UPB units, not a multiple of 32; N = 33000: 258 units): a warp folds BPI consecutive
// blocks per iteration in ceil(UPB/32) coalesced steps -- the last step of a block is ragged, which costs idle lanes but
// no bandwidth -- with BPI independent loads per lane and step in flight (twice that with the unrolled step loop).
// The mask unit of a step comes from L1 (the same UPB units for every block).
// ---------------------------------------------------------------------------------------
template <typename VT, int BPI>
__global__ void __launch_bounds__(kDecThreads, 4)
decrypt_count_rows_kernel(const VT *__restrict__ V, const uint64_t T, const uint32_t UPB, const VT *__restrict__ M,
                          uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    const uint32_t lane = threadIdx.x & 31u;
    pdl_enter();
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_groups = (T + BPI - 1) / BPI;
    const uint32_t steps = (UPB + 31u) >> 5;
    uint64_t my_count = 0;
    for (uint64_t grp = warp_global; grp < n_groups; grp += n_warps) {
        const uint64_t blk0 = grp * BPI;
        const VT *row = V + blk0 * UPB + lane;
        bool f[BPI];
#pragma unroll
        for (int b = 0; b < BPI; ++b) f[b] = false;
#pragma unroll(8 / BPI)
        for (uint32_t s = 0; s < steps; ++s) {
            const uint32_t u = s * 32u + lane;
            const bool in = u < UPB;
            const VT m = in ? __ldg(M + u) : vzero<VT>();
            VT v[BPI];
#pragma unroll
            for (int b = 0; b < BPI; ++b)
                v[b] = (in && blk0 + b < T) ? ld_stream(row + (uint64_t)b * UPB + s * 32u) : vzero<VT>();
#pragma unroll
            for (int b = 0; b < BPI; ++b) f[b] |= unit_fails(v[b], m);
        }
#pragma unroll
        for (int b = 0; b < BPI; ++b) {
            const bool bad = __any_sync(0xffffffffu, f[b]);
            if (lane == 0 && !bad && blk0 + b < T) ++my_count;
        }
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// ---------------------------------------------------------------------------------------
// Odd L (half of all contexts): a block is not a whole number of 16-byte units, but TWO consecutive blocks are -- a
// "double block" of 2L words = L units, 16-byte aligned whenever the ciphertext is.  A warp folds BPI double blocks per
// iteration exactly as the rows kernel does, with two verdicts per double block: a unit's low word belongs to the first
// block when its word index 2u is below L, its high word when 2u+1 is (the middle unit straddles the two blocks).  The
// key-mask words of a unit come from L1.  An odd last block is folded by the grid's first warp with 8-byte loads.
// ---------------------------------------------------------------------------------------
template <int BPI>
__global__ void __launch_bounds__(kDecThreads, 4)
decrypt_count_pairs_kernel(const uint4 *__restrict__ V4, const uint64_t T, const uint32_t L, const uint64_t *__restrict__ M,
                           uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    const uint32_t lane = threadIdx.x & 31u;
    pdl_enter();
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint64_t n_pairs = T >> 1;
    const uint64_t n_groups = (n_pairs + BPI - 1) / BPI;
    const uint32_t steps = (L + 31u) >> 5;
    uint64_t my_count = 0;
    for (uint64_t grp = warp_global; grp < n_groups; grp += n_warps) {
        const uint64_t pair0 = grp * BPI;
        const uint4 *row = V4 + pair0 * L + lane;
        bool f0[BPI], f1[BPI];
#pragma unroll
        for (int b = 0; b < BPI; ++b) f0[b] = f1[b] = false;
#pragma unroll(8 / BPI)
        for (uint32_t s = 0; s < steps; ++s) {
            const uint32_t u = s * 32u + lane;
            const bool in = u < L;
            const bool lo_first = 2u * u < L, hi_first = 2u * u + 1u < L;
            const uint64_t m_lo = in ? __ldg(M + (lo_first ? 2u * u : 2u * u - L)) : 0ull;
            const uint64_t m_hi = in ? __ldg(M + (hi_first ? 2u * u + 1u : 2u * u + 1u - L)) : 0ull;
            uint4 v[BPI];
#pragma unroll
            for (int b = 0; b < BPI; ++b)
                v[b] = (in && pair0 + b < n_pairs) ? ld_stream(row + (uint64_t)b * L + s * 32u) : vzero<uint4>();
#pragma unroll
            for (int b = 0; b < BPI; ++b) {
                const uint64_t lo = (uint64_t)v[b].x | ((uint64_t)v[b].y << 32), hi = (uint64_t)v[b].z | ((uint64_t)v[b].w << 32);
                const bool fl = (~lo & m_lo) != 0ull, fh = (~hi & m_hi) != 0ull;
                f0[b] |= (lo_first && fl) || (hi_first && fh);
                f1[b] |= (!lo_first && fl) || (!hi_first && fh);
            }
        }
#pragma unroll
        for (int b = 0; b < BPI; ++b) {
            const bool bad0 = __any_sync(0xffffffffu, f0[b]), bad1 = __any_sync(0xffffffffu, f1[b]);
            if (lane == 0 && pair0 + b < n_pairs) my_count += (bad0 ? 0u : 1u) + (bad1 ? 0u : 1u);
        }
    }
    if ((T & 1ull) && warp_global == 0) {
        const uint64_t *last = reinterpret_cast<const uint64_t *>(V4) + (T - 1) * L;
        bool f = false;
        for (uint32_t w = lane; w < L; w += 32) f |= (~__ldcs(last + w) & __ldg(M + w)) != 0ull;
        const bool bad = __any_sync(0xffffffffu, f);
        if (lane == 0 && !bad) ++my_count;
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// ---------------------------------------------------------------------------------------
// window: blocks of ANY length (odd L above all, and the even L the kernels above serve badly) on full 16-byte loads.
// The blocks of an odd-L ciphertext straddle 16-byte units, but the stream as a whole does not care: every warp owns one
// contiguous run of double blocks (so the run starts on a 16-byte boundary), split evenly over the grid's warps, and walks
// it in steps of 32 units = 64 words with every lane loading.
// A step's verdicts are two ballots -- Flo: the low words (window words 0,2,4..), Fhi: the high words (1,3,5..) -- and
// the block structure is laid over them afterwards: with r = words of the open block consumed before this window, the
// blocks that END inside the window end at word offsets (L-r) + j*L <= 64; lane j checks the bits of block j in the
// two ballots (plus, for j = 0, the carry: whether the open block failed in an earlier window), and the carry for the
// next window is what is set behind the last end.  r repeats with period L steps, so the bit masks of every (step, end)
// and the carry masks of every step are tabulated in shared memory once per CTA (before the PDL wait): a step is one
// load, one 16-byte mask unit, two votes, one table entry per lane, one carry entry per warp and a dozen logic ops per
// 512 bytes (31 SASS instructions) -- no fail strings, no idle lanes, one ragged step per warp per launch.  The loads
// form a software pipeline (UNROLL in flight per lane, refilled GROUP at a time as registers free up).
// ---------------------------------------------------------------------------------------
constexpr uint32_t kWindowMaxWords = 999;    // tables + mask units in shared memory: 48 L + 16 bytes (L > 64) within 48 KB

__device__ __forceinline__ uint4 lds128(const uint32_t addr) {                   // 16 bytes of shared memory by address
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}

__host__ __device__ __forceinline__ uint32_t bits_below(const uint32_t n) {      // n in 0..32: the n low bits
    return n >= 32u ? 0xffffffffu : (1u << n) - 1u;
}

template <int UNROLL, int MINB, int GROUP>
__global__ void __launch_bounds__(kDecThreads, MINB)
decrypt_count_window_kernel(const uint4 *__restrict__ V4, const uint64_t T, const uint32_t L, const uint64_t *__restrict__ M,
                            const uint64_t per, const uint32_t extra, uint64_t *scratch, uint64_t *count_out,
                            const __grid_constant__ PeerPush pp) {
    // per, extra: T/2 double blocks = per * (warps of the grid) + extra, divided on the host (a 64-bit division is
    // several hundred cycles of dependent instructions at the head of every warp)
    extern __shared__ uint4 smem_raw[];
    const uint32_t E = 64u / L + 1u;                                 // at most E blocks end inside one window
    uint4 *sC = smem_raw;                                            // [L]     carry: (lo mask, hi mask, keep, ends)
    uint4 *sE = sC + L;                                              // [L * E] end j of step s: (lo mask, hi mask, force, carry use)
    uint4 *sM = sE + (size_t)L * E;                                  // [L]     unit t of a double block: words 2t, 2t+1 (mod L)
    if (threadIdx.x == 0) sM[L] = make_uint4(0u, 0u, 0xffffffffu, 0u);      // [1]     the end that never counts
    // the key mask is on its way from global memory while the tables are built (they depend on L only) ...
    constexpr uint32_t kMaskRounds = (kWindowMaxWords + kDecThreads - 1) / kDecThreads;
    uint64_t m0[kMaskRounds], m1[kMaskRounds];
#pragma unroll
    for (uint32_t k = 0; k < kMaskRounds; ++k) {
        const uint32_t t = threadIdx.x + k * kDecThreads;
        if (t < L) {
            const uint32_t w0 = 2u * t < L ? 2u * t : 2u * t - L, w1 = w0 + 1u < L ? w0 + 1u : 0u;
            m0[k] = M[w0];
            m1[k] = M[w1];
        }
    }
    for (uint32_t i = threadIdx.x; i < L * E; i += blockDim.x) {
        const uint32_t s = i / E, j = i - s * E;
        const uint32_t r = (s * 64u) % L;                            // words of the open block before window s
        const uint32_t n_ends = (64u + r) / L;
        uint4 en = make_uint4(0u, 0u, 0xffffffffu, 0u);              // no such end in this window: never counted
        if (j < n_ends) {
            const uint32_t e = (L - r) + j * L, a = j * L > r ? j * L - r : 0u;      // words [a, e) of the window
            en = make_uint4(bits_below((e + 1u) >> 1) & ~bits_below((a + 1u) >> 1), bits_below(e >> 1) & ~bits_below(a >> 1),
                            0u, j == 0u ? 0xffffffffu : 0u);
        }
        sE[i] = en;
        if (j == 0u) {
            uint4 c = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u);         // no end: the open block stays open
            if (n_ends) {
                const uint32_t a2 = (L - r) + (n_ends - 1u) * L;                     // the next open block starts here
                c = make_uint4(~bits_below((a2 + 1u) >> 1), ~bits_below(a2 >> 1), 0u, n_ends);
            }
            sC[s] = c;
        }
    }
    pdl_enter();

    // ... and so are the first units of this warp's run while the mask goes to shared memory
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t pair0 = warp_global * per + min(warp_global, (uint64_t)extra);
    const uint64_t my_pairs = per + (warp_global < extra ? 1u : 0u);
    const uint32_t t0 = lane % L, inc32 = 32u % L;                   // this lane's unit of the double block, per step
    const uint32_t max_sub = (1u << 30) / L;                         // unit offsets inside a run stay below 2^30
    uint32_t cnt = 0;                                                // per lane: < 2^32 blocks per launch and lane
    uint64_t done = 0;
    uint32_t np = (uint32_t)min(my_pairs, (uint64_t)max_sub);
    const uint4 *src = V4 + pair0 * L + lane;
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = u * 32u + lane < np * L ? ld_stream(src + u * 32u) : vzero<uint4>();
#pragma unroll
    for (uint32_t k = 0; k < kMaskRounds; ++k) {
        const uint32_t t = threadIdx.x + k * kDecThreads;
        if (t < L) sM[t] = make_uint4((uint32_t)m0[k], (uint32_t)(m0[k] >> 32), (uint32_t)m1[k], (uint32_t)(m1[k] >> 32));
    }
    __syncthreads();

    // one step: v = this lane's unit of the window; CHECK: the run's last steps, where blocks past its end must not count.
    // The three table reads go through 32-bit shared-memory addresses that are advanced (and wrapped) by byte offsets.
#define CSGN_WINDOW_STEP(v, CHECK)                                                                                    \
    {                                                                                                                 \
        const uint4 m = lds128(ma);                                                                                   \
        const uint32_t Flo = __ballot_sync(0xffffffffu, ((~(v).x & m.x) | (~(v).y & m.y)) != 0u);                     \
        const uint32_t Fhi = __ballot_sync(0xffffffffu, ((~(v).z & m.z) | (~(v).w & m.w)) != 0u);                     \
        const uint4 c = lds128(ca);                                                                                   \
        const uint4 en = lds128(ea);                                                                                  \
        const uint32_t bad = (Flo & en.x) | (Fhi & en.y) | en.z | (carry & en.w);                                     \
        if (CHECK) {                                                                                                  \
            cnt += (bad == 0u && cb + lane < nblk) ? 1u : 0u;                                                         \
            cb += c.w;                                                                                                \
        } else {                                                                                                      \
            cnt += bad == 0u ? 1u : 0u;                                                                               \
        }                                                                                                             \
        carry = (carry & c.z) | (Flo & c.x) | (Fhi & c.y);                                                            \
        ma += m_inc;                                                                                                  \
        if (ma >= m_end) ma -= m_len;                                                                                 \
        ca += 16u;                                                                                                    \
        ea += e_inc;                                                                                                  \
        if (ca == c_end) {                                                                                            \
            ca = c_0;                                                                                                 \
            ea = e_0;                                                                                                 \
        }                                                                                                             \
    }

    const uint32_t m_len = 16u * L, m_inc = 16u * inc32, m_0 = (uint32_t)__cvta_generic_to_shared(sM + t0);
    const uint32_t m_end = (uint32_t)__cvta_generic_to_shared(sM) + m_len;
    const uint32_t c_0 = (uint32_t)__cvta_generic_to_shared(sC), c_end = c_0 + 16u * L;
    // lanes without an end of their own (lane >= E) stay on one entry that never counts
    const uint32_t e_0 = (uint32_t)__cvta_generic_to_shared(lane < E ? sE + lane : sM + L), e_inc = lane < E ? 16u * E : 0u;
    while (np) {
        const uint32_t nblk = 2u * np, n_units = np * L, steps = (n_units + 31u) >> 5, full_steps = n_units >> 5;
        uint32_t ma = m_0, ca = c_0, ea = e_0;
        uint32_t carry = 0, s0 = 0, cb = 0;
        // UNROLL loads stay in flight all the time: a unit's register is refilled with the unit UNROLL steps ahead as soon
        // as it has been folded (with loads issued in batches the memory pipe drains while a batch is being folded)
        for (; s0 + 2u * UNROLL <= full_steps; s0 += UNROLL) {       // every step folded and every step loaded here is whole
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                CSGN_WINDOW_STEP(v[u], false)
                v[u] = ld_stream(src + (UNROLL + u) * 32u);
                // no thread needs this barrier; it keeps the refills HERE, GROUP of them back to back (ptxas sinks them
                // to the end of the round, where the memory pipe has drained)
                if (u % GROUP == GROUP - 1) __syncwarp();
            }
            src += UNROLL * 32u;
        }
        cb = (s0 * 64u) / L;                                         // blocks that ended before the remaining steps
        for (; s0 < steps; s0 += UNROLL) {
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (s0 + u < steps) CSGN_WINDOW_STEP(v[u], true)     // warp-uniform
                v[u] = (s0 + UNROLL + u) * 32u + lane < n_units ? ld_stream(src + (UNROLL + u) * 32u) : vzero<uint4>();
                if (u % GROUP == GROUP - 1) __syncwarp();            // the last refills, too, go out as early as they can
            }
            src += UNROLL * 32u;
        }
        done += np;                                                  // (a run of more than 2^30 units goes on: next piece)
        np = (uint32_t)min(my_pairs - done, (uint64_t)max_sub);
        src = V4 + (pair0 + done) * L + lane;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = u * 32u + lane < np * L ? ld_stream(src + u * 32u) : vzero<uint4>();
    }
#undef CSGN_WINDOW_STEP
    uint64_t my_count = cnt;
    if ((T & 1ull) && warp_global == 0) {
        const uint64_t *last = reinterpret_cast<const uint64_t *>(V4) + (T - 1) * L;
        bool f = false;
        for (uint32_t w = lane; w < L; w += 32) f |= (~__ldcs(last + w) & __ldg(M + w)) != 0ull;
        const bool bad = __any_sync(0xffffffffu, f);
        if (lane == 0 && !bad) ++my_count;
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// Blocks longer than kDecMaxUnits units (N > 65536): one warp per block, 64-bit loads.
__global__ void __launch_bounds__(kDecThreads)
decrypt_count_generic_kernel(const uint64_t *__restrict__ V, const uint64_t T, const uint32_t L,
                             const uint64_t *__restrict__ M, uint64_t *scratch, uint64_t *count_out,
                             const __grid_constant__ PeerPush pp) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    uint64_t my_count = 0;
    pdl_enter();
    for (uint64_t blk = warp_global; blk < T; blk += n_warps) {
        const uint64_t *row = V + blk * L;
        bool f = false;
        for (uint32_t w = lane; w < L; w += 32) {
            const uint64_t m = __ldg(M + w);
            f |= ((~__ldcs(row + w)) & m) != 0ull;
        }
        const bool bad = __any_sync(0xffffffffu, f);
        if (lane == 0 && !bad) ++my_count;
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

// Persistent grid: exactly the CTAs that are resident at once (a partial second wave
// costs far more than it balances), never more than there is work for.
template <typename Kernel>
uint32_t resident_grid(Kernel kernel, size_t smem, uint64_t work_ctas) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kDecThreads, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    per_sm = (int)std::min<long>(per_sm, env_long("CSGN_DEC_CTAS_PER_SM", per_sm));
    const uint64_t cap = (uint64_t)device_props().sm_count * (uint64_t)std::max(per_sm, 1);
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(work_ctas, cap));
}

// One persistent wave of resident CTAs is best for a fold running alone at sizes where ramp and tail matter
// (160 MB: 26.4 vs 27.4 us); several shorter waves are better beyond a GiB (+3 %) and when the caller overlaps folds
// of several streams (back-fill behind the other stream's kernel: 24.5 -> 23.5 us per fold in the bench).
uint64_t fold_waves(uint64_t bytes, bool overlapped) {
    const bool many = overlapped || bytes >= (1ull << 30);
    return (uint64_t)std::max<long>(1, env_long("CSGN_DEC_WAVES", many ? 4 : 1));
}

template <typename VT, int UPBC, int UNROLL, int MINB>
cudaError_t launch_string(const void *v, uint64_t T, uint32_t upb, const void *mask, uint64_t *scratch, uint64_t *count_out,
                          const PeerPush &pp, cudaStream_t stream) {
    const uint64_t n_chunks = (T + 31) / 32;
    const size_t smem = (size_t)2 * upb * sizeof(VT) + (size_t)kDecWarps * upb * sizeof(uint32_t);
    const uint32_t cpw = (uint32_t)env_long("CSGN_DEC_CPW", 0);
    uint32_t grid;
    if (cpw)
        grid = (uint32_t)std::max<uint64_t>(1, (n_chunks + (uint64_t)kDecWarps * cpw - 1) / ((uint64_t)kDecWarps * cpw));
    else
        grid = resident_grid(decrypt_count_kernel<VT, UPBC, UNROLL, MINB>, smem, (n_chunks + kDecWarps - 1) / kDecWarps);
    return launch_kernel(decrypt_count_kernel<VT, UPBC, UNROLL, MINB>, grid, kDecThreads, smem, stream,
                         static_cast<const VT *>(v), T, upb, static_cast<const VT *>(mask), cpw, scratch, count_out, pp);
}

template <int UNROLL, int MINB>
cudaError_t launch_string_pairs(const uint64_t *v, uint64_t T, uint32_t L, const uint64_t *mask, uint64_t *scratch,
                                uint64_t *count_out, const PeerPush &pp, cudaStream_t stream) {
    const uint64_t n_chunks = std::max<uint64_t>(1, (T / 2 + 31) / 32);
    const size_t smem = (size_t)2 * L * sizeof(uint4) + (size_t)kDecWarps * 2 * L * sizeof(uint32_t);
    const uint32_t grid = resident_grid(decrypt_count_string_pairs_kernel<UNROLL, MINB>, smem, (n_chunks + kDecWarps - 1) / kDecWarps);
    return launch_kernel(decrypt_count_string_pairs_kernel<UNROLL, MINB>, grid, kDecThreads, smem, stream,
                         reinterpret_cast<const uint4 *>(v), T, L, mask, scratch, count_out, pp);
}

template <typename VT, int UPB, int U, int MINB>
cudaError_t launch_lanes(const void *v, uint64_t T, const uint64_t *host_mask, uint64_t *scratch, uint64_t *count_out,
                         const PeerPush &pp, bool overlapped, cudaStream_t stream) {
    ParamMask pm;
    memset(&pm, 0, sizeof pm);
    memcpy(&pm, host_mask, (size_t)UPB * sizeof(VT));
    constexpr uint32_t BPS = 32u / UPB;
    const uint64_t n_steps = (T + BPS - 1) / BPS;
    const uint64_t work_ctas = (n_steps + (uint64_t)kDecWarps * U - 1) / ((uint64_t)kDecWarps * U);
    uint32_t grid = resident_grid(decrypt_count_lanes_kernel<VT, UPB, U, MINB>, 0, work_ctas);
    grid = (uint32_t)std::min<uint64_t>(work_ctas, (uint64_t)grid * fold_waves(T * UPB * sizeof(VT), overlapped));
    return launch_kernel(decrypt_count_lanes_kernel<VT, UPB, U, MINB>, grid, kDecThreads, 0, stream,
                         static_cast<const VT *>(v), T, pm, scratch, count_out, pp);
}

template <int UPL, int BPI>
cudaError_t launch_wide(const uint64_t *v, uint64_t T, const uint64_t *mask, uint64_t *scratch,
                        uint64_t *count_out, const PeerPush &pp, bool overlapped, cudaStream_t stream) {
    const uint64_t n_groups = (T + BPI - 1) / BPI;
    const uint64_t work_ctas = (n_groups + kDecWarps - 1) / kDecWarps;
    uint32_t grid = resident_grid(decrypt_count_wide_kernel<UPL, BPI>, 0, work_ctas);
    grid = (uint32_t)std::min<uint64_t>(work_ctas, (uint64_t)grid * fold_waves(T * (uint64_t)UPL * 512u, overlapped));
    return launch_kernel(decrypt_count_wide_kernel<UPL, BPI>, grid, kDecThreads, 0, stream,
                         reinterpret_cast<const uint4 *>(v), T, reinterpret_cast<const uint4 *>(mask), scratch, count_out, pp);
}

template <typename VT, int BPI>
cudaError_t launch_rows(const uint64_t *v, uint64_t T, uint32_t upb, const uint64_t *mask, uint64_t *scratch,
                        uint64_t *count_out, const PeerPush &pp, bool overlapped, cudaStream_t stream) {
    const uint64_t n_groups = (T + BPI - 1) / BPI;
    const uint64_t work_ctas = (n_groups + kDecWarps - 1) / kDecWarps;
    uint32_t grid = resident_grid(decrypt_count_rows_kernel<VT, BPI>, 0, work_ctas);
    grid = (uint32_t)std::min<uint64_t>(work_ctas, (uint64_t)grid * fold_waves(T * (uint64_t)upb * sizeof(VT), overlapped));
    return launch_kernel(decrypt_count_rows_kernel<VT, BPI>, grid, kDecThreads, 0, stream, reinterpret_cast<const VT *>(v), T,
                         upb, reinterpret_cast<const VT *>(mask), scratch, count_out, pp);
}

template <int BPI>
cudaError_t launch_pairs(const uint64_t *v, uint64_t T, uint32_t L, const uint64_t *mask, uint64_t *scratch,
                         uint64_t *count_out, const PeerPush &pp, bool overlapped, cudaStream_t stream) {
    const uint64_t n_groups = std::max<uint64_t>(1, (T / 2 + BPI - 1) / BPI);
    const uint64_t work_ctas = (n_groups + kDecWarps - 1) / kDecWarps;
    uint32_t grid = resident_grid(decrypt_count_pairs_kernel<BPI>, 0, work_ctas);
    grid = (uint32_t)std::min<uint64_t>(work_ctas, (uint64_t)grid * fold_waves(T * (uint64_t)L * 8u, overlapped));
    return launch_kernel(decrypt_count_pairs_kernel<BPI>, grid, kDecThreads, 0, stream, reinterpret_cast<const uint4 *>(v), T, L,
                         mask, scratch, count_out, pp);
}


template <int UNROLL, int MINB, int GROUP>
cudaError_t launch_window(const uint64_t *v, uint64_t T, uint32_t L, const uint64_t *mask, uint64_t *scratch,
                          uint64_t *count_out, const PeerPush &pp, bool overlapped, cudaStream_t stream) {
    // a warp's run is worth at least one unrolled round of steps
    const uint64_t pairs_per_warp = std::max<uint64_t>(1, ((uint64_t)UNROLL * 32u + L - 1) / L);
    const uint64_t work_ctas = std::max<uint64_t>(1, (T / 2 + kDecWarps * pairs_per_warp - 1) / (kDecWarps * pairs_per_warp));
    const size_t ends = 64u / L + 1u;
    const size_t smem = ((size_t)L * (2 + ends) + 1) * sizeof(uint4);
    uint32_t grid = resident_grid(decrypt_count_window_kernel<UNROLL, MINB, GROUP>, smem, work_ctas);
    grid = (uint32_t)std::min<uint64_t>(work_ctas, (uint64_t)grid * fold_waves(T * (uint64_t)L * 8u, overlapped));
    const uint64_t n_warps = (uint64_t)grid * kDecWarps;
    return launch_kernel(decrypt_count_window_kernel<UNROLL, MINB, GROUP>, grid, kDecThreads, smem, stream,
                         reinterpret_cast<const uint4 *>(v), T, L, mask, (T / 2) / n_warps, (uint32_t)((T / 2) % n_warps), scratch,
                         count_out, pp);
}

// Push and/or publish + collect without a fold (an empty local shard still owes its peers a
// word; a collect may also be issued on its own).  One CTA.
__global__ void __launch_bounds__(kDecThreads)
peer_exchange_kernel(const __grid_constant__ PeerPush pp, const int do_push, const uint64_t value, uint64_t *count_out) {
    pdl_enter();
    if (do_push && threadIdx.x == 0) {
        if (count_out) *count_out = value;
        pp.local_ring[peer_slot(pp.seq)] = value;
    }
    if (pp.publish_n | pp.collect_n) {
        __syncthreads();
        peer_publish_collect(pp);
    }
}

#ifdef CSGN_BUILD_VARIANTS
// ---------------------------------------------------------------------------------------
// Bulk-ring variant of the same fold (small blocks, L4 <= 16, e.g. N=1247).
//
// One persistent CTA per SM.  A producer warp streams the ciphertext into a ring of
// shared-memory stages with 1-D bulk async copies (cp.async.bulk -> the TMA engine,
// completion counted in bytes on an mbarrier); a stage is kRingWarps chunks of 32 blocks.
// Consumer warp w folds chunk w of each stage out of shared memory exactly as above
// (ballot -> fail string -> one lane per block) and releases the stage through a second
// mbarrier.  The amount of data in flight is the ring (here ~160 KB per SM), not what the
// register file can hold, and HBM sees long sequential bursts.
// ---------------------------------------------------------------------------------------
constexpr int kRingWarps = 8;     // consumer warps = chunks per stage
constexpr int kRingThreads = (kRingWarps + 1) * 32;

template <int L4C, int STAGES>
__global__ void __launch_bounds__(kRingThreads, 1)
decrypt_count_ring_kernel(const uint4 *__restrict__ V4, const uint64_t T, const uint32_t L4rt,
                          const uint4 *__restrict__ M4, const __grid_constant__ ParamMask pmask,
                          uint64_t *scratch, uint64_t *count_out, const __grid_constant__ PeerPush pp) {
    extern __shared__ __align__(128) unsigned char ring_raw[];
    const uint32_t L4 = L4C ? (uint32_t)L4C : L4rt;
    const uint32_t chunk_units = 32u * L4;                       // uint4 per 32-block chunk
    const uint32_t stage_units = chunk_units * kRingWarps;
    uint4 *ring = reinterpret_cast<uint4 *>(ring_raw);           // STAGES * stage_units
    uint4 *sM2 = ring + (size_t)STAGES * stage_units;            // mask twice over
    uint32_t *sF = reinterpret_cast<uint32_t *>(sM2 + 2 * L4);   // fail strings, L4 words per warp
    uint64_t *full = reinterpret_cast<uint64_t *>(sF + kRingWarps * L4 + (((kRingWarps * L4) & 1u) ? 1u : 0u));
    uint64_t *empty = full + STAGES;

    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    pdl_enter();
    if (M4 == nullptr) {
        for (uint32_t i = threadIdx.x; i < 2 * L4; i += blockDim.x) sM2[i] = pmask.u[i < L4 ? i : i - L4];
    } else {
        for (uint32_t i = threadIdx.x; i < 2 * L4; i += blockDim.x) sM2[i] = M4[i < L4 ? i : i - L4];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, kRingWarps);
        }
        mbar_init_fence();
    }
    __syncthreads();

    const uint64_t n_units = T * L4;
    const uint64_t n_stage_units = (n_units + stage_units - 1) / stage_units;   // stage-sized pieces of the stream
    uint64_t my_count = 0;

    if (warp == kRingWarps) {
        // ---- producer: one lane keeps the ring full ------------------------------------
        if (lane == 0) {
            uint32_t slot = 0, phase = 0;
            for (uint64_t u = blockIdx.x; u < n_stage_units; u += gridDim.x) {
                mbar_wait(empty + slot, phase ^ 1u);             // consumers are done with this slot
                const uint64_t first = u * stage_units;
                const uint32_t units = (uint32_t)min((uint64_t)stage_units, n_units - first);
                mbar_expect_tx(full + slot, units * 16u);
                bulk_g2s(ring + (size_t)slot * stage_units, V4 + first, units * 16u, full + slot);
                if (++slot == STAGES) { slot = 0; phase ^= 1u; }
            }
        }
    } else {
        // ---- consumers: warp w folds chunk w of every stage ---------------------------
        uint32_t *sFw = sF + warp * L4;
        const uint4 *mk = sM2 + (lane % L4);
        const uint32_t step = 32u % L4;
        uint32_t slot = 0, phase = 0;
        for (uint64_t u = blockIdx.x; u < n_stage_units; u += gridDim.x) {
            mbar_wait(full + slot, phase);
            const uint4 *src = ring + (size_t)slot * stage_units + warp * chunk_units + lane;
            uint32_t koff = 0;
#pragma unroll
            for (uint32_t r = 0; r < (L4C ? (uint32_t)L4C : 16u); ++r) {
                if (L4C || r < L4) {
                    const uint4 v = src[32u * r];
                    const uint32_t bal = __ballot_sync(0xffffffffu, unit_fails(v, mk[koff]));
                    if (lane == 0) sFw[r] = bal;
                    koff += step;
                    if (koff >= L4) koff -= L4;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + slot);            // this warp's reads of the slot are done
            const uint64_t blk = (u * kRingWarps + warp) * 32u + lane;
            const uint32_t lo = lane * L4, hi = lo + L4;
            uint32_t any = 0;
            for (uint32_t w = lo >> 5; w <= (hi - 1) >> 5; ++w) {
                const uint32_t first = max(lo, w << 5) - (w << 5);
                const uint32_t last = min(hi, (w + 1) << 5) - (w << 5);
                const uint32_t m = (last - first == 32u) ? 0xffffffffu : (((1u << (last - first)) - 1u) << first);
                any |= sFw[w] & m;
            }
            my_count += (blk < T && any == 0u) ? 1u : 0u;
            __syncwarp();
            if (++slot == STAGES) { slot = 0; phase ^= 1u; }
        }
    }
    fold_and_publish(my_count, scratch, count_out, pp);
}

template <int L4C, int STAGES>
cudaError_t launch_ring(const uint64_t *v, uint64_t T, uint32_t L4, const uint64_t *mask, const uint64_t *host_mask,
                        uint64_t *scratch, uint64_t *count_out, const PeerPush &pp, cudaStream_t stream) {
    ParamMask pm;
    memset(&pm, 0, sizeof pm);
    const bool by_param = host_mask && L4 <= (uint32_t)kParamMaskUnits;
    if (by_param) memcpy(&pm, host_mask, (size_t)L4 * sizeof(uint4));
    const size_t stage_bytes = (size_t)32 * L4 * kRingWarps * sizeof(uint4);
    const size_t smem = STAGES * stage_bytes + 2 * L4 * sizeof(uint4) + ((size_t)kRingWarps * L4 + 1) * sizeof(uint32_t) +
                        2 * STAGES * sizeof(uint64_t) + 16;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(decrypt_count_ring_kernel<L4C, STAGES>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    const uint64_t n_stage_units = (T * L4 + (uint64_t)32 * L4 * kRingWarps - 1) / ((uint64_t)32 * L4 * kRingWarps);
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(n_stage_units, device_props().sm_count));
    return launch_kernel(decrypt_count_ring_kernel<L4C, STAGES>, grid, kRingThreads, smem, stream,
                         reinterpret_cast<const uint4 *>(v), T, L4,
                         by_param ? nullptr : reinterpret_cast<const uint4 *>(mask), pm, scratch, count_out, pp);
}


#endif  // CSGN_BUILD_VARIANTS

}  // namespace

cudaError_t launch_peer_exchange(const PeerPush &pp, bool do_push, uint64_t value, uint64_t *count_out,
                                 cudaStream_t stream) {
    count_launch();
    return launch_kernel(peer_exchange_kernel, 1, kDecThreads, 0, stream, pp, do_push ? 1 : 0, value, count_out);
}

bool build_has_variants() {
#ifdef CSGN_BUILD_VARIANTS
    return true;
#else
    return false;
#endif
}

cudaError_t launch_decrypt_count(const uint64_t *v, uint64_t T, uint32_t L, const uint64_t *mask,
                                 const uint64_t *host_mask, uint64_t *scratch, uint64_t *count_out,
                                 cudaStream_t stream, const PeerPush *peer, bool overlapped) {
    PeerPush pp;
    if (peer) pp = *peer;
    else memset(&pp, 0, sizeof pp);
    if (T == 0 || L == 0) {
        if (pp.world) return launch_peer_exchange(pp, true, 0, count_out, stream);
        return count_out ? cudaMemsetAsync(count_out, 0, sizeof(uint64_t), stream) : cudaSuccess;
    }
    const DeviceProps &dp = device_props();
    const bool aligned16 = ((reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(mask)) & 15u) == 0;
    const bool units16 = !(L & 1u) && aligned16;
    const uint32_t upb = units16 ? L / 2 : L;          // units per block: 16-byte or 8-byte
    cudaError_t err = cudaErrorInvalidValue;
    bool done = false;
#ifdef CSGN_BUILD_VARIANTS
    const long variant = env_long("CSGN_DEC_VARIANT", 0);
    if (units16 && upb == 10 && variant > 0) {
        done = true;
        switch (variant) {
            case 5: err = launch_ring<10, 4>(v, T, upb, mask, host_mask, scratch, count_out, pp, stream); break;
            case 6: err = launch_ring<10, 3>(v, T, upb, mask, host_mask, scratch, count_out, pp, stream); break;
            case 7: err = launch_ring<10, 5>(v, T, upb, mask, host_mask, scratch, count_out, pp, stream); break;
            case 8: err = launch_lanes<uint4, 10, 10, 4>(v, T, host_mask, scratch, count_out, pp, overlapped, stream); break;
            case 9: err = launch_lanes<uint4, 10, 8, 4>(v, T, host_mask, scratch, count_out, pp, overlapped, stream); break;
            case 11: err = launch_lanes<uint4, 10, 6, 6>(v, T, host_mask, scratch, count_out, pp, overlapped, stream); break;
            case 12: err = launch_lanes<uint4, 10, 16, 3>(v, T, host_mask, scratch, count_out, pp, overlapped, stream); break;
            case 1: err = launch_string<uint4, 10, 10, 4>(v, T, upb, mask, scratch, count_out, pp, stream); break;
            case 2: err = launch_string<uint4, 10, 5, 5>(v, T, upb, mask, scratch, count_out, pp, stream); break;
            case 3: err = launch_string<uint4, 10, 5, 6>(v, T, upb, mask, scratch, count_out, pp, stream); break;
            case 4: err = launch_string<uint4, 10, 2, 8>(v, T, upb, mask, scratch, count_out, pp, stream); break;
            case 13: err = launch_string<uint4, 10, 10, 3>(v, T, upb, mask, scratch, count_out, pp, stream); break;
            default: done = false;
        }
    }
#endif
    const bool force_string = env_long("CSGN_DEC_STRING", 0) != 0;
    // the older odd-L kernels stay reachable (tests, tools/oddl_probe.py) through their own switches
    const bool older_odd = env_long("CSGN_DEC_PAIRS_MIN", -1) >= 0 || env_long("CSGN_DEC_STRING_PAIRS", -1) >= 0 ||
                           env_long("CSGN_DEC_GENERIC", 0) || force_string || !env_long("CSGN_DEC_WINDOW", 1);
    if (done) {
    } else if (L >= 3 && L <= kWindowMaxWords && !older_odd && (reinterpret_cast<uintptr_t>(v) & 15u) == 0 &&
               (env_long("CSGN_DEC_WINDOW_ALL", 0) ||
                ((L & 1u) ? (L == 3 || L >= 17)
                          : (units16 && upb > 16 && upb < 129 && upb % 32 != 0 && env_long("CSGN_DEC_WINDOW_EVEN", 1))))) {
        // the window walk over the flat stream on full 16-byte loads: every odd block length from 17 words up and 3 words
        // (B200, tools/window_probe.py, fraction of the copy peak at 160 MB against the kernels before it: L = 3 0.90 /
        // 0.68, 19 0.88 / 0.75, 33 0.88 / 0.73, 65 0.89 / 0.65, 97 0.88 / 0.66, 193 0.88 / 0.70; 1.6 GB at L = 65: 1.03 /
        // 0.77; odd L of 5..15 words stay with the lane-aligned 8-byte kernel: 0.91-0.95 / 0.86-0.88), and the even
        // lengths whose unit count fits neither the lane-aligned nor the warp-per-block kernels (17..128 units, not a
        // multiple of 32).  Forms <loads in flight, CTAs per SM, refills per group>: tools/window_waves_probe.py.
        const long form = env_long("CSGN_DEC_WINDOW", 1);
        // fewer, deeper warps (<12,2,1>) from half a GiB on and for long blocks, whose runs are few double blocks per warp
        // (160 MB: L = 641 0.84 / 0.80, L = 999 0.80 / 0.75; 640 MB: 1.00 / 0.96 over L = 129 .. 999)
        const bool big = (uint64_t)T * L * 8u >= (1ull << 29) || L >= 300;
        switch (form > 1 ? form : big ? 4 : 11) {
#define CSGN_WINDOW_CASE(ID, U, B, G) \
    case ID: err = launch_window<U, B, G>(v, T, L, mask, scratch, count_out, pp, overlapped, stream); break;
            CSGN_WINDOW_CASE(2, 6, 4, 1) CSGN_WINDOW_CASE(4, 12, 2, 1) CSGN_WINDOW_CASE(6, 8, 3, 2) CSGN_WINDOW_CASE(11, 6, 4, 2)
            default: err = launch_window<8, 3, 1>(v, T, L, mask, scratch, count_out, pp, overlapped, stream); break;
#undef CSGN_WINDOW_CASE
        }
    } else if ((L & 1u) && L >= (uint32_t)env_long("CSGN_DEC_PAIRS_MIN", 49) && !force_string &&
               !env_long("CSGN_DEC_GENERIC", 0) && (reinterpret_cast<uintptr_t>(v) & 15u) == 0) {
        // odd L of 49 words and more: double blocks of 2L words folded warp-per-double-block on 16-byte loads (B200,
        // tools/oddl_probe.py, fraction of the copy peak against the 8-byte-unit kernel: L = 65 0.65 / 0.64, L = 97
        // 0.66 / 0.58, L = 193 0.70 / 0.59); shorter odd blocks take the fail-string walk over double blocks below
        // (L = 19 0.75 / 0.65, L = 33 0.73 / 0.70)
        const uint32_t steps = (L + 31u) / 32u;
        const long bpi = env_long("CSGN_DEC_ROWS_BPI", steps >= 8 ? 1 : steps >= 4 ? 2 : 4);
        if (bpi >= 4) err = launch_pairs<4>(v, T, L, mask, scratch, count_out, pp, overlapped, stream);
        else if (bpi >= 2) err = launch_pairs<2>(v, T, L, mask, scratch, count_out, pp, overlapped, stream);
        else err = launch_pairs<1>(v, T, L, mask, scratch, count_out, pp, overlapped, stream);
    } else if ((L & 1u) && L >= 17 && L <= kDecMaxUnits && !env_long("CSGN_DEC_GENERIC", 0) && env_long("CSGN_DEC_STRING_PAIRS", 1) &&
               (reinterpret_cast<uintptr_t>(v) & 15u) == 0) {
        // odd L of 17..48 words: the fail-string walk on 16-byte units over double blocks
        err = launch_string_pairs<4, 4>(v, T, L, mask, scratch, count_out, pp, stream);
    } else if (upb > kDecMaxUnits || env_long("CSGN_DEC_GENERIC", 0)) {
        const uint64_t want = (T + kDecWarps - 1) / kDecWarps;
        const uint64_t ctas_per_sm = (uint64_t)env_long("CSGN_DEC_CTAS_PER_SM", 8);
        const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(want, dp.sm_count * ctas_per_sm));
        err = launch_kernel(decrypt_count_generic_kernel, grid, kDecThreads, 0, stream, v, T, L, mask, scratch, count_out, pp);
    } else if (upb <= 16 && host_mask && !force_string) {
        // every short block shape has its own lane-aligned instantiation (no cliff between contexts)
#define CSGN_LANES_CASE(VT, UPB, U) \
    case UPB: err = launch_lanes<VT, UPB, U, 3>(v, T, host_mask, scratch, count_out, pp, overlapped, stream); break;
        if (units16) {
            switch (upb) {
                CSGN_LANES_CASE(uint4, 1, 12) CSGN_LANES_CASE(uint4, 2, 12) CSGN_LANES_CASE(uint4, 3, 12)
                CSGN_LANES_CASE(uint4, 4, 12) CSGN_LANES_CASE(uint4, 5, 12) CSGN_LANES_CASE(uint4, 6, 12)
                CSGN_LANES_CASE(uint4, 7, 12) CSGN_LANES_CASE(uint4, 8, 12) CSGN_LANES_CASE(uint4, 9, 12)
                CSGN_LANES_CASE(uint4, 10, 12) CSGN_LANES_CASE(uint4, 11, 12) CSGN_LANES_CASE(uint4, 12, 12)
                CSGN_LANES_CASE(uint4, 13, 12) CSGN_LANES_CASE(uint4, 14, 12) CSGN_LANES_CASE(uint4, 15, 12)
                CSGN_LANES_CASE(uint4, 16, 12)
            }
        } else {
            switch (upb) {
                CSGN_LANES_CASE(uint2, 1, 16) CSGN_LANES_CASE(uint2, 2, 16) CSGN_LANES_CASE(uint2, 3, 16)
                CSGN_LANES_CASE(uint2, 4, 16) CSGN_LANES_CASE(uint2, 5, 16) CSGN_LANES_CASE(uint2, 6, 16)
                CSGN_LANES_CASE(uint2, 7, 16) CSGN_LANES_CASE(uint2, 8, 16) CSGN_LANES_CASE(uint2, 9, 16)
                CSGN_LANES_CASE(uint2, 10, 16) CSGN_LANES_CASE(uint2, 11, 16) CSGN_LANES_CASE(uint2, 12, 16)
                CSGN_LANES_CASE(uint2, 13, 16) CSGN_LANES_CASE(uint2, 14, 16) CSGN_LANES_CASE(uint2, 15, 16)
                CSGN_LANES_CASE(uint2, 16, 16)
            }
        }
#undef CSGN_LANES_CASE
    } else if (units16 && upb % 32 == 0 && upb <= 256 && !force_string && env_long("CSGN_DEC_WIDE", 1)) {
        switch (upb / 32) {
            case 1: err = launch_wide<1, 8>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            case 2: err = launch_wide<2, 4>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            case 3: err = launch_wide<3, 3>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            case 4: err = launch_wide<4, 2>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;   // N=16383
            case 5: err = launch_wide<5, 2>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            case 6: err = launch_wide<6, 2>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            case 7: err = launch_wide<7, 1>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
            default: err = launch_wide<8, 1>(v, T, mask, scratch, count_out, pp, overlapped, stream); break;
        }
    } else if (units16 && upb >= (uint32_t)env_long("CSGN_DEC_ROWS_MIN", 129) && !force_string) {
        // long blocks whose unit count is not a multiple of 32 (N = 33000: 258 units = 9 steps).  A lane keeps
        // (blocks per iteration) x (unrolled steps) = 8 loads in flight; with 8 or more steps per block one block per
        // iteration is enough and splits the work finest (B200: 0.40 -> 0.94 of the copy peak at 165 MB, 1.09 at 1.6 GB).
        // (Shorter blocks and 8-byte units -- odd L -- stay with the fail-string kernel, which measured faster there.)
        const uint32_t steps = (upb + 31u) / 32u;
        const long bpi = env_long("CSGN_DEC_ROWS_BPI", steps >= 8 ? 1 : steps >= 4 ? 2 : 4);
        if (bpi >= 4) err = launch_rows<uint4, 4>(v, T, upb, mask, scratch, count_out, pp, overlapped, stream);
        else if (bpi >= 2) err = launch_rows<uint4, 2>(v, T, upb, mask, scratch, count_out, pp, overlapped, stream);
        else err = launch_rows<uint4, 1>(v, T, upb, mask, scratch, count_out, pp, overlapped, stream);
    } else if (units16) {
        err = launch_string<uint4, 0, 4, 4>(v, T, upb, mask, scratch, count_out, pp, stream);
    } else {
        err = launch_string<uint2, 0, 8, 4>(v, T, upb, mask, scratch, count_out, pp, stream);
    }
    count_launch();
    return err;
}

}  // namespace csgn
