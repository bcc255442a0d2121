// capi_internal.cuh -- shared by the translation units that implement include/csgn.h (capi.cu: library, buffers,
// the hot path, batches; capi_comm.cu: the peer communicator; capi_io.cu: save / load): the handle types, the
// process-wide state and the small helpers every entry point uses.  Internal: not part of the C ABI.
#pragma once

#include "../../include/csgn.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"

// ---------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------
struct csgn_buf {
    uint64_t *d = nullptr;     // device words, n_blocks * L valid
    uint64_t n_blocks = 0;
    uint32_t L = 0;
    uint64_t cap_words = 0;    // allocated words (>= n_blocks*L); 0 for views
    bool owns = true;
    bool recycle = false;      // storage of an upload: goes back to the upload cache, not to the pool
    mutable cudaEvent_t ready = nullptr;  // an upload on the copy stream still in flight
    mutable cudaStream_t last_stream = nullptr;  // the stream of the last operation that touched the words
};

struct csgn_key {
    uint64_t *d_positions = nullptr;  // D secret positions (for batched encryption)
    uint64_t *d_mask = nullptr;  // L words
    std::vector<uint64_t> h_mask;  // the same, host side (small masks ride in kernel parameters)
    uint64_t N = 0;
    uint32_t L = 0, D = 0;
};

struct csgn_perm {
    uint32_t *d_map = nullptr;   // N entries: (src_word << 6) | right_shift
    uint32_t *d_slice_map = nullptr;  // 64*L entries for the bit-sliced kernel (permute.cu), or null
    uint64_t N = 0;
    uint32_t L = 0;
};

struct csgn_comm {
    int rank = 0, world = 1;
    uint64_t *box_local = nullptr;                 // this rank's mailbox (cudaMalloc: exportable)
    uint64_t *box[csgn::kPeerMaxWorld] = {};       // every rank's mailbox as mapped here
    bool ipc_opened[csgn::kPeerMaxWorld] = {};
    bool connected = false;
    uint64_t seq = 0;                              // sequence number of the next push
    uint64_t published = 0;                        // pushes [0, published) have been stored to the peers
    uint64_t *d_local_ring = nullptr;              // kPeerRing words: this rank's counts by slot
    uint64_t *d_status = nullptr;                  // [0] timeout flag, [1] blocking-call total
    uint64_t timeout_ns = 30ull * 1000 * 1000 * 1000;
    std::string rendezvous_file;                   // written by csgn_comm_connect_dir, removed on free
};

namespace csgn {
namespace detail {

struct State {
    bool inited = false;
    int device = -1;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D uploads run here, overlapping kernels of the work stream
    // Batch entry points (csgn_*_batch) spread their independent items over these streams, forked from and joined
    // back into the current work stream, so that the tail of one kernel overlaps the ramp of the next item's.
    static constexpr int kMaxLanes = 4;
    cudaStream_t lane[kMaxLanes] = {};
    cudaEvent_t lane_done[kMaxLanes] = {};
    cudaEvent_t fork_point = nullptr;
    int n_lanes = 2;
    std::vector<cudaEvent_t> event_pool;
    // Storage of freed uploads, each with the event that marks its last use: an upload takes a slot whose event has
    // COMPLETED, so its copy can start at once on the copy stream and no allocation (which across streams may go to the
    // driver, milliseconds) sits on the steady-state path of "upload operands, multiply, decrypt, free".
    struct UploadSlot {
        uint64_t *d;
        uint64_t cap_words;
        cudaEvent_t freed;
    };
    std::vector<UploadSlot> upload_cache;
    uint64_t upload_cache_words = 0;
    uint64_t *d_scratch = nullptr;   // [2] blocking-call result, [4..6] checksum, [8 + 2k, 9 + 2k] fold scratch of launch k mod 64
    uint32_t fold_slot = 0;
    uint64_t *h_result = nullptr;    // pinned, 8 words
};
extern State g;
extern unsigned g_launches_since_switch;   // launches since the caller last changed streams (see streams_alternate)

// Record the message of a failure for csgn_last_error() and return `code`.
int fail(int code, const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define CU(call)                                            \
    do {                                                    \
        cudaError_t e_ = (call);                            \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call); \
    } while (0)

#define NEED_INIT()                                                                              \
    do {                                                                                         \
        if (!g.inited) return fail(CSGN_ERR_NOT_INITIALIZED, "csgn_init has not been called");   \
        int cur_ = -1;                                                                           \
        if (cudaGetDevice(&cur_) != cudaSuccess || cur_ != g.device) CU(cudaSetDevice(g.device)); \
    } while (0)

int dev_alloc(uint64_t words, uint64_t **out, cudaStream_t stream = nullptr);   // stream-ordered, on the work stream by default
void dev_free(void *p);
cudaEvent_t take_event();
// Called for every operand of every operation: orders the work stream after a pending upload of `b`
// and remembers which stream touched the words last.
void await_upload(const csgn_buf *b);
int new_buf(uint64_t n_blocks, uint32_t L, uint64_t cap_words, csgn_buf **out, cudaStream_t stream = nullptr);
// The next pair of fold scratch words (running count, CTA ticket); every launch gets its own.
uint64_t *next_fold_scratch();

// Scope of one batch call: item i runs with the work stream set to way i % n, where way 0 is the caller's own stream
// and ways 1.. are the library's side lanes, forked from the caller's stream on construction; join() makes the
// caller's stream wait for every side lane and restores it.
struct LaneScope {
    cudaStream_t home;
    int used = 1;
    bool active;
    explicit LaneScope(uint32_t n_items) : home(g.stream), active(g.n_lanes > 1 && n_items > 1) {
        if (!active) return;
        used = (int)std::min<uint32_t>(n_items, (uint32_t)g.n_lanes);
        cudaEventRecord(g.fork_point, home);
        for (int i = 1; i < used; ++i) cudaStreamWaitEvent(g.lane[i], g.fork_point, 0);
    }
    void enter(uint32_t item) {
        if (!active) return;
        const uint32_t way = item % (uint32_t)used;
        g.stream = way == 0 ? home : g.lane[way];
        g_launches_since_switch = 0;              // the items of a batch overlap: the launchers' multi-wave forms apply
    }
    void join() {
        if (!active) return;
        g.stream = home;
        for (int i = 1; i < used; ++i) {
            cudaEventRecord(g.lane_done[i], g.lane[i]);
            cudaStreamWaitEvent(home, g.lane_done[i], 0);
        }
        active = false;
    }
    ~LaneScope() { join(); }
};

}  // namespace detail
}  // namespace csgn
