// fold.cuh -- the pieces of the decrypt fold that more than one kernel uses: the unit predicate, the
// key mask carried in kernel parameters, and the CTA -> grid -> (peers) reduction of the satisfied-block
// count.  Included by decrypt.cu (K3) and by mul.cu (the fused multiply -> fold kernels).
//
//   block k is satisfied  <=>  for every word w: (v[k*L+w] & M[w]) == M[w]     (reference src/SecretKey.cpp:131-137)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "peer.cuh"

namespace csgn {

// The key mask of a small block travels in the kernel parameters: after a multiply has streamed through L2 a
// 160-byte mask in global memory is a DRAM miss at the head of every CTA, while the parameter bank is always hot.
constexpr int kParamMaskUnits = 16;  // up to 32 words per block (N <= 2048)
struct ParamMask {
    uint4 u[kParamMaskUnits];
};

#ifdef __CUDACC__
// some key bit inside this 16-byte (8-byte) unit is zero
__device__ __forceinline__ bool unit_fails(const uint4 v, const uint4 m) {
    return (((~v.x) & m.x) | ((~v.y) & m.y) | ((~v.z) & m.z) | ((~v.w) & m.w)) != 0u;
}
__device__ __forceinline__ bool unit_fails(const uint2 v, const uint2 m) {
    return (((~v.x) & m.x) | ((~v.y) & m.y)) != 0u;
}

// Grid-level end of a fold.  Thread 0 of every CTA of the grid adds (1 << 40 | cta_count) to the launch's scratch
// word with ONE atomic: the low 40 bits accumulate the count, the high 24 bits are the CTA ticket, and the value the
// atomic returns tells the last CTA both that it is last and what the total is -- no second atomic, no fence (the
// total is taken from the atomic's own result, not from memory another CTA wrote).  The last CTA re-arms the scratch
// word, writes the total, and -- with a PeerPush (sharded decrypt) -- leaves it in the rank's local ring; if this
// launch closes a batch it then publishes the batch to every rank's mailbox over NVLink and collects the requested
// totals (peer.cuh).  Counts are < 2^40 and grids < 2^24 CTAs (launchers cap both).
// Every thread of every CTA must call this exactly once; `cta_count` is read from thread 0 only.
__device__ __forceinline__ void grid_publish(const uint64_t cta_count, uint64_t *scratch, uint64_t *count_out,
                                             const PeerPush &pp) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        const unsigned long long old =
            atomicAdd(reinterpret_cast<unsigned long long *>(scratch), (1ull << 40) | (unsigned long long)cta_count);
        const bool last = (old >> 40) == (unsigned long long)gridDim.x - 1ull;
        if (last) {
            const uint64_t total = (old + cta_count) & kPeerCountMask;
            *scratch = 0;                    // every other CTA's atomic has been performed: re-arm for the next launch
            if (count_out) *count_out = total;
            if (pp.world) pp.local_ring[peer_slot(pp.seq)] = total;
        }
        s_last = last ? 1 : 0;
    }
    if (pp.world && (pp.publish_n | pp.collect_n)) {      // grid-uniform: the barrier is not divergent
        __syncthreads();
        if (s_last) peer_publish_collect(pp);
    }
}

// Lane counts -> warp (shuffles) -> CTA (shared memory) -> grid_publish.  The CTA size must be a multiple of 32.
__device__ __forceinline__ void fold_and_publish(uint64_t lane_count, uint64_t *scratch, uint64_t *count_out,
                                                 const PeerPush &pp) {
    __shared__ uint64_t s_warp[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) lane_count += __shfl_xor_sync(0xffffffffu, lane_count, off);
    if (lane == 0) s_warp[warp] = lane_count;
    __syncthreads();
    uint64_t cta = 0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int w = 0; w < nw; ++w) cta += s_warp[w];
    }
    grid_publish(cta, scratch, count_out, pp);
}
#endif

}  // namespace csgn
