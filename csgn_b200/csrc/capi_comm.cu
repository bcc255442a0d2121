// capi_comm.cu -- csgn_comm_*: per-rank mailboxes mapped over NVLink, and the sharded decrypt whose fold kernel
// does the cross-GPU exchange itself (csrc/peer.cuh has the device side and the protocol).
#include "capi_internal.cuh"

#include <errno.h>
#include <signal.h>
#include <unistd.h>

#include <sys/stat.h>
#include <unistd.h>

#include <ctime>

using namespace csgn;
using namespace csgn::detail;

extern "C" {

// ---------------------------------------------------------------------------
// sharded decrypt: fold + cross-GPU exchange in one kernel (peer.cuh)
// ---------------------------------------------------------------------------
namespace {
constexpr size_t kMailboxBytes = (size_t)kPeerRing * kPeerMaxWorld * sizeof(uint64_t);

int comm_ready(const csgn_comm *comm) {
    if (!comm) return fail(CSGN_ERR_INVALID_ARGUMENT, "null communicator");
    if (!comm->connected)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "communicator of %d ranks is not connected (csgn_comm_connect)", comm->world);
    return CSGN_OK;
}

// Parameters of a launch that pushes (with_push) and, when collect_n > 0, closes the batch:
// publishes everything unpublished and collects pushes last-lag-collect_n+1 .. last-lag.
// How a batch of sharded folds is closed: 1 = by a small exchange kernel of its own after all n items have run on
// the lanes; 0 = by the last item's kernel (its last CTA publishes and collects), which then runs after the join,
// alone.  Measured on 2 and 8 B200 (profiles/README.md); CSGN_PEER_CLOSE_KERNEL overrides under CSGN_TUNING.
constexpr long kCloseWithOwnKernel = 1;

int fill_push(const csgn_comm *comm, bool with_push, uint32_t collect_n, uint32_t lag, uint64_t *totals, PeerPush *pp) {
    memset(pp, 0, sizeof *pp);
    const uint64_t last = with_push ? comm->seq : comm->seq - 1;       // most recent push after this launch
    const uint64_t issued = last + 1;
    const uint64_t unpublished = issued - comm->published;
    if (collect_n) {
        if (!totals) return fail(CSGN_ERR_INVALID_ARGUMENT, "collect without a destination");
        if ((uint64_t)collect_n + lag > kPeerMaxPending)
            return fail(CSGN_ERR_INVALID_ARGUMENT, "collect window of %u pushes trailing by %u exceeds %u", collect_n, lag,
                        kPeerMaxPending);
        if ((uint64_t)collect_n + lag > issued)
            return fail(CSGN_ERR_INVALID_ARGUMENT, "collect of %u pushes trailing by %u, only %llu issued so far", collect_n,
                        lag, (unsigned long long)issued);
    } else if (unpublished > kPeerMaxPending) {
        return fail(CSGN_ERR_INVALID_ARGUMENT, "more than %u pushes without a collect", kPeerMaxPending);
    }
    for (int q = 0; q < comm->world; ++q) pp->box[q] = comm->box[q];
    pp->local_ring = comm->d_local_ring;
    pp->seq = last;
    pp->totals = totals;
    pp->status = comm->d_status;
    pp->timeout_ns = comm->timeout_ns;
    pp->world = (uint32_t)comm->world;
    pp->rank = (uint32_t)comm->rank;
    pp->publish_n = collect_n ? (uint32_t)unpublished : 0u;
    pp->collect_n = collect_n;
    pp->collect_lag = lag;
    return CSGN_OK;
}
}  // namespace

int csgn_comm_create(int rank, int world, csgn_comm **out, unsigned char *handle_out) {
    NEED_INIT();
    if (!out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output handle");
    if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "bad communicator shape: rank %d of %d (at most %d ranks)", rank, world,
                    kPeerMaxWorld);
    csgn_comm *c = new csgn_comm;
    c->rank = rank;
    c->world = world;
    const char *t = std::getenv("CSGN_PEER_TIMEOUT_MS");
    if (t && *t) c->timeout_ns = (uint64_t)std::max(1L, std::atol(t)) * 1000000ull;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&c->box_local), kMailboxBytes);
    if (e == cudaSuccess) e = cudaMemset(c->box_local, 0, kMailboxBytes);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&c->d_status), 4 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemset(c->d_status, 0, 4 * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&c->d_local_ring), kPeerRing * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMemset(c->d_local_ring, 0, kPeerRing * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaDeviceSynchronize();   // the mailbox is zero before any peer can learn of it
    if (e == cudaSuccess && handle_out) {
        static_assert(sizeof(cudaIpcMemHandle_t) == CSGN_IPC_HANDLE_BYTES, "IPC handle size");
        cudaIpcMemHandle_t h;
        memset(&h, 0, sizeof h);
        if (world > 1) e = cudaIpcGetMemHandle(&h, c->box_local);
        if (e == cudaSuccess) memcpy(handle_out, &h, sizeof h);
    }
    if (e != cudaSuccess) {
        if (c->box_local) cudaFree(c->box_local);
        if (c->d_status) cudaFree(c->d_status);
        if (c->d_local_ring) cudaFree(c->d_local_ring);
        delete c;
        return cuda_fail(e, "communicator mailbox");
    }
    c->box[rank] = c->box_local;
    c->connected = (world == 1);
    *out = c;
    return CSGN_OK;
}

int csgn_comm_connect(csgn_comm *comm, const unsigned char *handles) {
    NEED_INIT();
    if (!comm || !handles) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    for (int q = 0; q < comm->world; ++q) {
        if (q == comm->rank || comm->box[q]) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)q * CSGN_IPC_HANDLE_BYTES, sizeof h);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(CSGN_ERR_CUDA, "cannot map the mailbox of rank %d over NVLink (cudaIpcOpenMemHandle: %s)", q,
                        cudaGetErrorString(e));
        }
        comm->box[q] = static_cast<uint64_t *>(p);
        comm->ipc_opened[q] = true;
    }
    comm->connected = true;
    return CSGN_OK;
}

int csgn_comm_connect_ptrs(csgn_comm *comm, void *const *peer_mailboxes) {
    NEED_INIT();
    if (!comm || !peer_mailboxes) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    for (int q = 0; q < comm->world; ++q) {
        if (q == comm->rank) continue;
        if (!peer_mailboxes[q] || (reinterpret_cast<uintptr_t>(peer_mailboxes[q]) & 7u))
            return fail(CSGN_ERR_INVALID_ARGUMENT, "mailbox pointer of rank %d is null or misaligned", q);
        comm->box[q] = static_cast<uint64_t *>(peer_mailboxes[q]);
    }
    comm->connected = true;
    return CSGN_OK;
}

int csgn_comm_connect_dir(csgn_comm *comm, const unsigned char *handle, const char *dir, const char *tag,
                          int timeout_ms) {
    NEED_INIT();
    if (!comm || !handle || !dir || !tag) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    if (comm->world == 1) {
        comm->connected = true;
        return CSGN_OK;
    }
    auto path_of = [&](int r) {
        return std::string(dir) + "/csgn_" + tag + "_" + std::to_string(comm->world) + "_" + std::to_string(r) + ".handle";
    };
    const time_t started = time(nullptr);
    const std::string mine = path_of(comm->rank), tmp = mine + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot write %s", tmp.c_str());
    // the file carries the writer's pid after the handle: a file left behind by a crashed job with the same tag names a
    // process that no longer exists and is ignored, whatever its age
    const uint64_t my_pid = (uint64_t)getpid();
    const bool ok = fwrite(handle, 1, CSGN_IPC_HANDLE_BYTES, f) == CSGN_IPC_HANDLE_BYTES && fwrite(&my_pid, sizeof my_pid, 1, f) == 1;
    if (fclose(f) != 0 || !ok || rename(tmp.c_str(), mine.c_str()) != 0)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot publish %s", mine.c_str());
    comm->rendezvous_file = mine;
    std::vector<unsigned char> all((size_t)comm->world * CSGN_IPC_HANDLE_BYTES, 0);
    memcpy(all.data() + (size_t)comm->rank * CSGN_IPC_HANDLE_BYTES, handle, CSGN_IPC_HANDLE_BYTES);
    for (int q = 0; q < comm->world; ++q) {
        if (q == comm->rank) continue;
        const std::string theirs = path_of(q);
        for (long waited_ms = 0;; waited_ms += 2) {
            struct stat st;
            if (stat(theirs.c_str(), &st) == 0 && st.st_size == CSGN_IPC_HANDLE_BYTES + (off_t)sizeof(uint64_t) &&
                st.st_mtime >= started - 120) {
                FILE *g2 = fopen(theirs.c_str(), "rb");
                uint64_t their_pid = 0;
                bool got = g2 && fread(all.data() + (size_t)q * CSGN_IPC_HANDLE_BYTES, 1, CSGN_IPC_HANDLE_BYTES, g2) ==
                                     CSGN_IPC_HANDLE_BYTES && fread(&their_pid, sizeof their_pid, 1, g2) == 1;
                if (g2) fclose(g2);
                // alive? (kill with signal 0 only probes; EPERM also means "exists")
                if (got && their_pid != 0 && (kill((pid_t)their_pid, 0) == 0 || errno == EPERM)) break;
            }
            if (waited_ms >= timeout_ms)
                return fail(CSGN_ERR_TIMEOUT, "rank %d of %d did not publish %s within %d ms", q, comm->world,
                            theirs.c_str(), timeout_ms);
            usleep(2000);
        }
    }
    return csgn_comm_connect(comm, all.data());
}

void *csgn_comm_mailbox(const csgn_comm *comm, size_t *bytes) {
    if (bytes) *bytes = kMailboxBytes;
    return comm ? comm->box_local : nullptr;
}

uint32_t csgn_comm_pending(const csgn_comm *comm) { return comm ? (uint32_t)(comm->seq - comm->published) : 0; }

void csgn_comm_slot_tag(uint64_t seq, uint32_t *slot, uint64_t *tag) {
    if (slot) *slot = peer_slot(seq);
    if (tag) *tag = peer_tag(seq);
}

int csgn_comm_free(csgn_comm *comm) {
    if (!comm) return CSGN_OK;
    if (!comm->rendezvous_file.empty()) remove(comm->rendezvous_file.c_str());
    if (g.inited) {
        cudaStreamSynchronize(g.stream);
        for (int q = 0; q < comm->world; ++q)
            if (comm->ipc_opened[q]) cudaIpcCloseMemHandle(comm->box[q]);
        if (comm->box_local) cudaFree(comm->box_local);
        if (comm->d_status) cudaFree(comm->d_status);
        if (comm->d_local_ring) cudaFree(comm->d_local_ring);
    }
    delete comm;
    return CSGN_OK;
}

int csgn_decrypt_sharded_async(const csgn_buf *c, const csgn_key *key, csgn_comm *comm, uint32_t collect_n,
                               uint32_t collect_lag, uint64_t *device_totals, uint64_t *device_local) {
    NEED_INIT();
    if (!c || !key) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    int rc = comm_ready(comm);
    if (rc != CSGN_OK) return rc;
    if (c->L != key->L)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "ciphertext has %u words per block, key expects %u", c->L, key->L);
    if (c->n_blocks > kPeerCountMask) return fail(CSGN_ERR_INVALID_ARGUMENT, "shard too large for a 40-bit count");
    rc = need_dense(c);          // the fold that publishes to the peers reads one array
    if (rc != CSGN_OK) return rc;
    PeerPush pp;
    rc = fill_push(comm, true, collect_n, collect_lag, device_totals, &pp);
    if (rc != CSGN_OK) return rc;
    acquire_read(c);
    cudaError_t e = launch_decrypt_count(c->d, c->n_blocks, c->L, key->d_mask,
                                         key->h_mask.empty() ? nullptr : key->h_mask.data(), fold_scratch(), device_local,
                                         g.stream, &pp, folds_overlap());
    if (e != cudaSuccess) return cuda_fail(e, "sharded decrypt kernel");
    comm->seq += 1;
    if (collect_n) comm->published = comm->seq;
    return CSGN_OK;
}

int csgn_mul_decrypt_sharded_async(const csgn_buf *a, const csgn_buf *b, const csgn_key *key, csgn_buf **out,
                                   csgn_comm *comm, uint32_t collect_n, uint32_t collect_lag, uint64_t *device_totals,
                                   uint64_t *device_local) {
    NEED_INIT();
    int rc = check_mul_operands(a, b);
    if (rc == CSGN_OK) rc = check_key(a, key);
    if (rc == CSGN_OK) rc = comm_ready(comm);
    if (rc != CSGN_OK) return rc;
    if (a->n_blocks * b->n_blocks > kPeerCountMask) return fail(CSGN_ERR_INVALID_ARGUMENT, "shard too large for a 40-bit count");
    PeerPush pp;
    rc = fill_push(comm, true, collect_n, collect_lag, device_totals, &pp);
    if (rc != CSGN_OK) return rc;
    csgn_buf *dst = nullptr;
    bool allocated = false;
    rc = fused_out(a, b, out, &dst, &allocated);
    if (rc != CSGN_OK) return rc;
    rc = enqueue_mul(a, b, dst, key, device_local, &pp);
    if (rc != CSGN_OK) {
        if (allocated) csgn_buf_free(dst);
        return rc;
    }
    if (allocated) *out = dst;
    comm->seq += 1;
    if (collect_n) comm->published = comm->seq;
    return CSGN_OK;
}

int csgn_mul_decrypt_sharded_batch_async(const csgn_buf *const *a, const csgn_buf *const *b, uint32_t n,
                                         const csgn_key *key, csgn_buf **out, csgn_comm *comm, uint32_t collect_lag,
                                         uint64_t *device_totals) {
    NEED_INIT();
    if (n == 0) return CSGN_OK;
    if (!a || !b || !device_totals) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    if ((uint64_t)n + collect_lag > kPeerMaxPending)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "batch of %u folds trailing by %u exceeds %u", n, collect_lag, kPeerMaxPending);
    if (env_long("CSGN_PEER_CLOSE_KERNEL", kCloseWithOwnKernel)) {
        // every item runs on the lanes, overlapped with its neighbours; the batch is then closed by ONE small kernel
        // on the caller's stream that publishes the n counts over NVLink and collects the sums
        {
            LaneScope lanes(n);
            for (uint32_t i = 0; i < n; ++i) {
                lanes.enter(i);
                int rc = csgn_mul_decrypt_sharded_async(a[i], b[i], key, out ? &out[i] : nullptr, comm, 0, 0, nullptr, nullptr);
                if (rc != CSGN_OK) return rc;
            }
        }
        return csgn_comm_collect_async(comm, n, collect_lag, device_totals);
    }
    {
        LaneScope lanes(n - 1);
        for (uint32_t i = 0; i + 1 < n; ++i) {
            lanes.enter(i);
            int rc = csgn_mul_decrypt_sharded_async(a[i], b[i], key, out ? &out[i] : nullptr, comm, 0, 0, nullptr, nullptr);
            if (rc != CSGN_OK) return rc;
        }
    }   // joined: the closing launch is ordered after every push it publishes
    return csgn_mul_decrypt_sharded_async(a[n - 1], b[n - 1], key, out ? &out[n - 1] : nullptr, comm, n, collect_lag,
                                          device_totals, nullptr);
}

int csgn_comm_collect_async(csgn_comm *comm, uint32_t n, uint32_t lag, uint64_t *device_totals) {
    NEED_INIT();
    int rc = comm_ready(comm);
    if (rc != CSGN_OK) return rc;
    if (n == 0) return CSGN_OK;
    if (comm->seq == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "collect before any push");
    PeerPush pp;
    rc = fill_push(comm, false, n, lag, device_totals, &pp);
    if (rc != CSGN_OK) return rc;
    cudaError_t e = launch_peer_exchange(pp, false, 0, nullptr, g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "collect kernel");
    comm->published = comm->seq;
    return CSGN_OK;
}

int csgn_decrypt_sharded(const csgn_buf *c, const csgn_key *key, csgn_comm *comm, uint8_t *bit, uint64_t *total) {
    NEED_INIT();
    if (!bit) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output");
    int rc = comm_ready(comm);
    if (rc != CSGN_OK) return rc;
    rc = csgn_decrypt_sharded_async(c, key, comm, 1, 0, comm->d_status + 1, nullptr);
    if (rc != CSGN_OK) return rc;
    CU(cudaMemcpyAsync(g.h_result, comm->d_status, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    note_synced(g.stream);
    if (g.h_result[0] != 0 || g.h_result[1] == UINT64_MAX) {
        cudaMemsetAsync(comm->d_status, 0, sizeof(uint64_t), g.stream);
        return fail(CSGN_ERR_TIMEOUT, "sharded decrypt: a peer's count did not arrive within %llu ms",
                    (unsigned long long)(comm->timeout_ns / 1000000ull));
    }
    *bit = (uint8_t)(g.h_result[1] & 1u);
    if (total) *total = g.h_result[1];
    return CSGN_OK;
}


int csgn_decrypt_sharded_batch_async(const csgn_buf *const *c, uint32_t n, const csgn_key *key, csgn_comm *comm,
                                     uint32_t collect_lag, uint64_t *device_totals) {
    NEED_INIT();
    if (n == 0) return CSGN_OK;
    if (!c || !device_totals) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    if ((uint64_t)n + collect_lag > kPeerMaxPending)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "batch of %u folds trailing by %u exceeds %u", n, collect_lag, kPeerMaxPending);
    if (env_long("CSGN_PEER_CLOSE_KERNEL", kCloseWithOwnKernel)) {
        {
            LaneScope lanes(n);
            for (uint32_t i = 0; i < n; ++i) {
                lanes.enter(i);
                int rc = csgn_decrypt_sharded_async(c[i], key, comm, 0, 0, nullptr, nullptr);
                if (rc != CSGN_OK) return rc;
            }
        }
        return csgn_comm_collect_async(comm, n, collect_lag, device_totals);
    }
    {
        LaneScope lanes(n - 1);
        for (uint32_t i = 0; i + 1 < n; ++i) {
            lanes.enter(i);
            int rc = csgn_decrypt_sharded_async(c[i], key, comm, 0, 0, nullptr, nullptr);
            if (rc != CSGN_OK) return rc;
        }
    }   // joined: the closing launch is ordered after every push it publishes
    return csgn_decrypt_sharded_async(c[n - 1], key, comm, n, collect_lag, device_totals, nullptr);
}

}  // extern "C"
