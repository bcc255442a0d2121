// peer.cuh -- the cross-GPU half of a sharded decrypt, done by the decrypt kernel itself.
//
// A ciphertext sharded by block range decrypts to the parity of the SUM of the
// per-shard satisfied-block counts (reference src/SecretKey.cpp:139 folds blocks with
// (dec + _dec) % 2, which is associative and commutative).  That sum is the only exchange step
// of the whole path, and it is one word per rank.  Instead of following the fold kernel with a
// library all-reduce, every fold kernel leaves its count in a local ring (push number `seq`),
// and the launch that closes a batch has its last CTA
//
//   publish  store (tag | count) of every not-yet-published push straight into the matching
//            slot of EVERY rank's mailbox -- peer device memory mapped over NVLink / NVSwitch
//            (8-byte st.relaxed.sys, one thread per (push, rank) pair, posted writes), and
//   collect  poll its OWN mailbox until the slots of the requested pushes carry every rank's
//            word with the expected tag, add them up and write the totals next to the caller's
//            other results.  The collected window may trail the published one (`lag`): a step
//            that collects the PREVIOUS step's batch never waits for a slower peer.
//
// Only the closing launch touches remote memory: a kernel that has stores in flight to a peer
// does not retire until NVLink has acknowledged them (~1.4 us measured per launch), so pushing
// from every fold kernel would tax each one of them.
//
// A mailbox is kPeerRing slots x kPeerMaxWorld words, zero-initialised; the tag
// (seq / ring) % (2^24-1) + 1 is never zero and differs between consecutive uses of a slot,
// so a word is valid exactly when its tag matches -- payload and flag travel in ONE 64-bit
// store, and no fence or second flag write is needed.  Counts are < 2^40 (a B200 holds fewer
// than 2^35 words).
//
// Protocol (same contract as any collective library): every rank issues the same sequence of
// pushes and collects.  A collect returns only when every rank has published the slots it
// covers; at most kPeerMaxPending pushes may stay unpublished and a collect window (n + lag)
// spans at most kPeerMaxPending pushes, so a rank can never lap a slot (ring = 4 x that) which
// a slower rank has still to read.  A peer that never arrives trips a timeout: the totals read
// UINT64_MAX and the status word is set -- the GPU is never left spinning.
#pragma once

#include <stdint.h>

namespace csgn {

constexpr int kPeerMaxWorld = 16;
constexpr uint32_t kPeerRing = 256;
constexpr uint32_t kPeerMaxPending = 64;
constexpr uint64_t kPeerCountMask = (1ull << 40) - 1;

struct PeerPush {
    uint64_t *box[kPeerMaxWorld];   // box[q]: rank q's mailbox as mapped into this process
    uint64_t *local_ring;           // kPeerRing words: this rank's own counts by slot
    uint64_t seq;                   // sequence number of this launch's push
    uint64_t *totals;               // collect_n sums go here (device memory), oldest first
    uint64_t *status;               // device word, set to 1 when a collect timed out
    uint64_t timeout_ns;
    uint32_t world;                 // 0: plain single-GPU launch, nothing below is touched
    uint32_t rank;
    uint32_t publish_n;             // > 0: the last CTA publishes pushes seq-publish_n+1 .. seq to every rank
    uint32_t collect_n;             // > 0: ... and collects pushes seq-lag-collect_n+1 .. seq-lag
    uint32_t collect_lag;
    uint32_t pad_;
};

inline __host__ __device__ uint32_t peer_slot(uint64_t seq) { return (uint32_t)(seq % kPeerRing); }
inline __host__ __device__ uint64_t peer_tag(uint64_t seq) { return (seq / kPeerRing) % 0xFFFFFFull + 1ull; }
inline __host__ __device__ uint64_t peer_word(uint64_t seq, uint64_t count) {
    return (peer_tag(seq) << 40) | (count & kPeerCountMask);
}

#ifdef __CUDACC__
__device__ __forceinline__ void peer_store(uint64_t *p, uint64_t v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t peer_load(const uint64_t *p) {
    uint64_t v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t global_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Whole CTA: publish this rank's counts of the most recent publish_n pushes to every rank
// (entry (j, q) by its own thread), then wait for pushes seq-lag-n+1 .. seq-lag of every rank and
// sum per push.  The caller has made local_ring[slot(seq)] visible to the CTA (barrier).
__device__ __forceinline__ void peer_publish_collect(const PeerPush &pp) {
    __shared__ unsigned long long s_tot[kPeerMaxPending];
    __shared__ int s_timed_out;
    const uint32_t n = pp.collect_n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s_tot[i] = 0ull;
    if (threadIdx.x == 0) s_timed_out = 0;
    for (uint32_t e = threadIdx.x; e < pp.publish_n * pp.world; e += blockDim.x) {
        const uint32_t j = e / pp.world, q = e - j * pp.world;
        const uint64_t s = pp.seq - (uint64_t)(pp.publish_n - 1u - j);
        const uint32_t slot = peer_slot(s);
        peer_store(pp.box[q] + slot * kPeerMaxWorld + pp.rank, peer_word(s, pp.local_ring[slot]));
    }
    __syncthreads();
    if (n == 0) return;
    const uint64_t *mine = pp.box[pp.rank];
    const uint64_t last = pp.seq - pp.collect_lag;
    const uint64_t t0 = global_ns();
    for (uint32_t e = threadIdx.x; e < n * pp.world; e += blockDim.x) {
        const uint32_t j = e / pp.world, q = e - j * pp.world;
        const uint64_t s = last - (uint64_t)(n - 1u - j);
        const uint64_t *src = mine + peer_slot(s) * kPeerMaxWorld + q;
        const uint64_t tag = peer_tag(s);
        uint64_t v = peer_load(src);
        while ((v >> 40) != tag) {
            if (global_ns() - t0 > pp.timeout_ns) {
                s_timed_out = 1;
                v = 0;
                break;
            }
            v = peer_load(src);
        }
        atomicAdd(&s_tot[j], (unsigned long long)(v & kPeerCountMask));
    }
    __syncthreads();
    const bool bad = s_timed_out != 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) pp.totals[i] = bad ? ~0ull : (uint64_t)s_tot[i];
    if (bad && threadIdx.x == 0) *pp.status = 1ull;
}
#endif

}  // namespace csgn
