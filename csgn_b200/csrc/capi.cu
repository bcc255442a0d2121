// capi.cu -- the C ABI of include/csgn.h: device state, buffer handles, and the
// thin argument checking in front of the kernel launchers.  No CPU compute path
// exists here: every operation either launches an sm_100a kernel or fails.
#include "capi_internal.cuh"

#include <atomic>
#include <cstdarg>

#define CSGN_VERSION_STRING "csgn-b200 0.1 (sm_100a)"

namespace csgn {
namespace detail {

State g;
unsigned g_launches_since_switch = 1u << 30;
DeviceProps g_props;
std::atomic<uint64_t> g_launches{0};
thread_local std::string t_error;

int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    t_error = buf;
    return code;
}

int cuda_fail(cudaError_t e, const char *what) {
    // leave the sticky-error state readable but do not hide it
    cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? CSGN_ERR_OUT_OF_MEMORY : CSGN_ERR_CUDA, "%s: %s (%s)", what,
                cudaGetErrorString(e), cudaGetErrorName(e));
}

// Every fold launch takes the next pair of scratch words (running count, CTA ticket) and leaves them
// zeroed, so decrypts enqueued on different streams (csgn_set_stream between calls) may run concurrently.
constexpr uint32_t kFoldSlots = 64;
uint64_t *next_fold_scratch() {
    uint64_t *p = g.d_scratch + 8 + 2 * (g.fold_slot % kFoldSlots);
    g.fold_slot += 1;
    return p;
}

int dev_alloc(uint64_t words, uint64_t **out, cudaStream_t stream) {
    *out = nullptr;
    if (words == 0) return CSGN_OK;
    void *p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, words * sizeof(uint64_t), stream ? stream : g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "cudaMallocAsync");
    *out = static_cast<uint64_t *>(p);
    return CSGN_OK;
}

void dev_free(void *p) {
    if (p) cudaFreeAsync(p, g.stream);
}

cudaEvent_t take_event() {
    if (!g.event_pool.empty()) {
        cudaEvent_t e = g.event_pool.back();
        g.event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    return e;
}

constexpr uint64_t kUploadCacheMaxWords = (64ull << 20) / 8;      // per buffer: larger uploads are not latency-bound
constexpr uint64_t kUploadCacheTotalWords = (1ull << 30) / 8;     // all slots together
constexpr size_t kUploadCacheSlots = 512;

// Storage for an upload of `words` words: a cached slot whose last use has completed, or null.
uint64_t *take_upload_slot(uint64_t words, uint64_t *cap_words) {
    size_t best = SIZE_MAX;
    for (size_t i = 0; i < g.upload_cache.size(); ++i) {
        const State::UploadSlot &sl = g.upload_cache[i];
        if (sl.cap_words < words || sl.cap_words > 2 * words + 64) continue;
        if (best != SIZE_MAX && g.upload_cache[best].cap_words <= sl.cap_words) continue;
        if (cudaEventQuery(sl.freed) != cudaSuccess) {
            cudaGetLastError();      // cudaErrorNotReady is not an error
            continue;
        }
        best = i;
    }
    if (best == SIZE_MAX) return nullptr;
    State::UploadSlot sl = g.upload_cache[best];
    g.upload_cache[best] = g.upload_cache.back();
    g.upload_cache.pop_back();
    g.upload_cache_words -= sl.cap_words;
    g.event_pool.push_back(sl.freed);
    *cap_words = sl.cap_words;
    return sl.d;
}

// Hand the storage of a freed upload to the cache (true), or decline (the caller frees it).
bool give_upload_slot(uint64_t *d, uint64_t cap_words) {
    if (!d || cap_words == 0 || cap_words > kUploadCacheMaxWords || g.upload_cache.size() >= kUploadCacheSlots ||
        g.upload_cache_words + cap_words > kUploadCacheTotalWords)
        return false;
    cudaEvent_t e = nullptr;
    if (!g.event_pool.empty()) {
        e = g.event_pool.back();
        g.event_pool.pop_back();
    } else if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (cudaEventRecord(e, g.stream) != cudaSuccess) {      // everything enqueued so far may still read the buffer
        cudaGetLastError();
        g.event_pool.push_back(e);
        return false;
    }
    g.upload_cache.push_back({d, cap_words, e});
    g.upload_cache_words += cap_words;
    return true;
}

// Called for every operand of every operation: order the work stream after a pending upload of `b`
// (first consumer only) and remember which stream touched the words last.
void await_upload(const csgn_buf *b) {
    if (!b) return;
    if (b->ready) {
        cudaStreamWaitEvent(g.stream, b->ready, 0);
        g.event_pool.push_back(b->ready);
        b->ready = nullptr;
    }
    b->last_stream = g.stream;
}

// Before storage is handed back (pool or upload cache) on the CURRENT stream: if the last operation on it ran on
// another stream (the caller multiplexes streams and frees later), the current stream first waits for that one.
void order_after_last_use(const csgn_buf *b) {
    if (!b->last_stream || b->last_stream == g.stream) return;
    cudaEvent_t e = take_event();
    if (!e) return;
    if (cudaEventRecord(e, b->last_stream) == cudaSuccess) cudaStreamWaitEvent(g.stream, e, 0);
    else cudaGetLastError();     // the caller destroyed that stream: its work was enqueued before, nothing to order
    g.event_pool.push_back(e);
}

int new_buf(uint64_t n_blocks, uint32_t L, uint64_t cap_words, csgn_buf **out, cudaStream_t stream) {
    if (L == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "words per block must be > 0");
    if (n_blocks > (UINT64_MAX / 8) / L) return fail(CSGN_ERR_INVALID_ARGUMENT, "block count overflows");
    csgn_buf *b = new csgn_buf;
    b->n_blocks = n_blocks;
    b->L = L;
    b->cap_words = std::max<uint64_t>(cap_words, n_blocks * L);
    int rc = dev_alloc(b->cap_words, &b->d, stream);
    if (rc != CSGN_OK) {
        delete b;
        return rc;
    }
    *out = b;
    return CSGN_OK;
}

}  // namespace detail

using namespace detail;

const DeviceProps &device_props() { return g_props; }
void set_device_props(const DeviceProps &p) { g_props = p; }
// A caller that moves between streams from one call to the next (csgn_set_stream) is enqueueing independent
// ciphertexts so that their kernels overlap; launchers that care (the decrypt fold) then prefer several shorter waves of
// CTAs, which back-fill behind another stream's kernel, over one persistent wave, which is best for a kernel running alone.
void count_launch(unsigned n) {
    g_launches.fetch_add(n, std::memory_order_relaxed);
    if (g_launches_since_switch < (1u << 30)) g_launches_since_switch += n;
}
bool streams_alternate() { return g_launches_since_switch < 8; }
uint64_t launches() { return g_launches.load(std::memory_order_relaxed); }

// Tuning knobs (CSGN_MUL_*, CSGN_DEC_*, CSGN_PERM_*, CSGN_PDL ...) are looked up only when
// CSGN_TUNING is set in the environment at csgn_init: a production launch does not pay a
// dozen getenv() scans (~4 us per launch), tests and tools/ opt in.
static bool g_tuning = true;   // until csgn_init has looked

long env_long(const char *name, long dflt) {
    if (!g_tuning) return dflt;
    const char *s = std::getenv(name);
    if (!s || !*s) return dflt;
    char *end = nullptr;
    long v = std::strtol(s, &end, 10);
    return (end && *end == 0) ? v : dflt;
}

}  // namespace csgn

using namespace csgn;
using namespace csgn::detail;

// ---------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------
extern "C" {

int csgn_init(int device) {
    csgn::g_tuning = true;
    if (device < 0) device = (int)env_long("CSGN_DEVICE", env_long("LOCAL_RANK", 0));
    csgn::g_tuning = std::getenv("CSGN_TUNING") != nullptr;
    if (g.inited) {
        if (g.device == device) return CSGN_OK;
        return fail(CSGN_ERR_INVALID_ARGUMENT, "already bound to device %d (one process per GPU)", g.device);
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(CSGN_ERR_NO_DEVICE, "no CUDA device visible (%s); this engine has no CPU path",
                    e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
    }
    if (device >= n) return fail(CSGN_ERR_NO_DEVICE, "device %d requested, %d visible", device, n);
    CU(cudaSetDevice(device));
    cudaDeviceProp p;
    CU(cudaGetDeviceProperties(&p, device));
    if (p.major != 10)
        return fail(CSGN_ERR_NO_DEVICE, "device %d is sm_%d%d; the kernels are built for sm_100a (B200) only",
                    device, p.major, p.minor);
    DeviceProps dp;
    dp.device = device;
    dp.sm_count = p.multiProcessorCount;
    dp.cc = p.major * 10 + p.minor;
    dp.smem_optin = p.sharedMemPerBlockOptin;
    set_device_props(dp);

    CU(cudaStreamCreateWithFlags(&g.own_stream, cudaStreamNonBlocking));
    g.stream = g.own_stream;
    CU(cudaStreamCreateWithFlags(&g.copy_stream, cudaStreamNonBlocking));
    {
        const char *e = std::getenv("CSGN_LANES");
        g.n_lanes = e && *e ? std::max(1, std::min((int)State::kMaxLanes, std::atoi(e))) : 2;
    }
    for (int i = 0; i < g.n_lanes; ++i) {
        CU(cudaStreamCreateWithFlags(&g.lane[i], cudaStreamNonBlocking));
        CU(cudaEventCreateWithFlags(&g.lane_done[i], cudaEventDisableTiming));
    }
    CU(cudaEventCreateWithFlags(&g.fork_point, cudaEventDisableTiming));
    // keep freed blocks in the pool: a*b chains reuse them without going to the driver
    cudaMemPool_t pool;
    CU(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    // Callers may enqueue independent ciphertexts on several streams (csgn_set_stream) so that kernels overlap.  An
    // allocation must then never make its stream wait for another stream's pending frees -- that would serialise the
    // streams again -- so the pool only reuses memory whose free has completed or is ordered by the caller's own events.
    int off = 0;
    CU(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off));
    CU(cudaMalloc(reinterpret_cast<void **>(&g.d_scratch), (8 + 2 * kFoldSlots) * sizeof(uint64_t)));
    CU(cudaMemset(g.d_scratch, 0, (8 + 2 * kFoldSlots) * sizeof(uint64_t)));
    CU(cudaHostAlloc(reinterpret_cast<void **>(&g.h_result), 8 * sizeof(uint64_t), cudaHostAllocDefault));
    g.device = device;
    g.inited = true;
    return CSGN_OK;
}

int csgn_shutdown(void) {
    if (!g.inited) return CSGN_OK;
    cudaSetDevice(g.device);
    cudaStreamSynchronize(g.copy_stream);
    cudaStreamSynchronize(g.stream);
    for (const State::UploadSlot &sl : g.upload_cache) {
        cudaFreeAsync(sl.d, g.stream);
        cudaEventDestroy(sl.freed);
    }
    cudaStreamSynchronize(g.stream);
    cudaFree(g.d_scratch);
    cudaFreeHost(g.h_result);
    for (cudaEvent_t e : g.event_pool) cudaEventDestroy(e);
    for (int i = 0; i < State::kMaxLanes; ++i) {
        if (g.lane[i]) {
            cudaStreamSynchronize(g.lane[i]);
            cudaStreamDestroy(g.lane[i]);
        }
        if (g.lane_done[i]) cudaEventDestroy(g.lane_done[i]);
    }
    if (g.fork_point) cudaEventDestroy(g.fork_point);
    cudaStreamDestroy(g.copy_stream);
    cudaStreamDestroy(g.own_stream);
    g = State();
    return CSGN_OK;
}

int csgn_is_initialized(void) { return g.inited ? 1 : 0; }
const char *csgn_last_error(void) { return t_error.c_str(); }
const char *csgn_version(void) { return CSGN_VERSION_STRING; }

int csgn_device_info(int *sm_count, uint64_t *hbm_total, uint64_t *hbm_free, int *cc) {
    NEED_INIT();
    size_t fr = 0, tot = 0;
    CU(cudaMemGetInfo(&fr, &tot));
    if (sm_count) *sm_count = device_props().sm_count;
    if (hbm_total) *hbm_total = tot;
    if (hbm_free) *hbm_free = fr;
    if (cc) *cc = device_props().cc;
    return CSGN_OK;
}

int csgn_set_stream(void *cuda_stream) {
    NEED_INIT();
    cudaStream_t next = cuda_stream ? static_cast<cudaStream_t>(cuda_stream) : g.own_stream;
    if (next != g.stream) g_launches_since_switch = 0;
    g.stream = next;
    return CSGN_OK;
}

void *csgn_get_stream(void) { return g.inited ? static_cast<void *>(g.stream) : nullptr; }

int csgn_sync(void) {
    NEED_INIT();
    CU(cudaStreamSynchronize(g.copy_stream));   // uploads whose buffers nobody has consumed yet
    CU(cudaStreamSynchronize(g.stream));
    return CSGN_OK;
}

uint64_t csgn_launch_count(void) { return launches(); }

uint32_t csgn_words_per_block(uint64_t N) { return (uint32_t)(N / 64 + (N % 64 ? 1 : 0)); }

int csgn_host_alloc(size_t bytes, void **out) {
    NEED_INIT();
    if (!out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output pointer");
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return CSGN_OK;
}

int csgn_host_free(void *p) {
    NEED_INIT();
    if (p) CU(cudaFreeHost(p));
    return CSGN_OK;
}

// ---------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------
int csgn_buf_alloc(uint64_t n_blocks, uint32_t L, csgn_buf **out) {
    NEED_INIT();
    if (!out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output handle");
    return new_buf(n_blocks, L, 0, out);
}

int csgn_buf_upload(const uint64_t *host_words, uint64_t n_blocks, uint32_t L, csgn_buf **out) {
    NEED_INIT();
    if (!out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output handle");
    if (n_blocks && !host_words) return fail(CSGN_ERR_INVALID_ARGUMENT, "null host words");
    // Allocation and copy are ordered on the copy stream, so an upload overlaps whatever the work
    // stream is running; the first consumer on the work stream waits for the `ready` event.
    csgn_buf *b = nullptr;
    uint64_t cap = 0;
    uint64_t *cached = (n_blocks && n_blocks <= kUploadCacheMaxWords / L) ? take_upload_slot(n_blocks * L, &cap) : nullptr;
    if (cached) {
        b = new csgn_buf;
        b->d = cached;
        b->n_blocks = n_blocks;
        b->L = L;
        b->cap_words = cap;
    } else {
        int rc = new_buf(n_blocks, L, 0, &b, g.copy_stream);
        if (rc != CSGN_OK) return rc;
    }
    b->recycle = true;
    if (n_blocks) {
        cudaError_t e = cudaMemcpyAsync(b->d, host_words, n_blocks * L * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                        g.copy_stream);
        if (e == cudaSuccess) {
            b->ready = take_event();
            e = b->ready ? cudaEventRecord(b->ready, g.copy_stream) : cudaStreamSynchronize(g.copy_stream);
        }
        if (e != cudaSuccess) {
            csgn_buf_free(b);
            return cuda_fail(e, "cudaMemcpyAsync(H2D)");
        }
    }
    *out = b;
    return CSGN_OK;
}

int csgn_buf_wrap(void *device_words, uint64_t n_blocks, uint32_t L, csgn_buf **out) {
    NEED_INIT();
    if (!out || L == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad view arguments");
    if (n_blocks && !device_words) return fail(CSGN_ERR_INVALID_ARGUMENT, "null device pointer");
    if (reinterpret_cast<uintptr_t>(device_words) & 7u)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "device pointer must be 8-byte aligned");
    csgn_buf *b = new csgn_buf;
    b->d = static_cast<uint64_t *>(device_words);
    b->n_blocks = n_blocks;
    b->L = L;
    b->cap_words = 0;
    b->owns = false;
    *out = b;
    return CSGN_OK;
}

int csgn_buf_clone(const csgn_buf *src, csgn_buf **out) {
    NEED_INIT();
    if (!src || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    csgn_buf *b = nullptr;
    int rc = new_buf(src->n_blocks, src->L, 0, &b);
    if (rc != CSGN_OK) return rc;
    await_upload(src);
    cudaError_t e = launch_concat(src->d, src->n_blocks * src->L, nullptr, 0, b->d, g.stream);
    if (e != cudaSuccess) {
        csgn_buf_free(b);
        return cuda_fail(e, "clone kernel");
    }
    *out = b;
    return CSGN_OK;
}

int csgn_buf_slice(const csgn_buf *src, uint64_t first_block, uint64_t n_blocks, csgn_buf **out) {
    NEED_INIT();
    if (!src || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (first_block > src->n_blocks || n_blocks > src->n_blocks - first_block)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "block range [%llu,+%llu) outside %llu blocks",
                    (unsigned long long)first_block, (unsigned long long)n_blocks, (unsigned long long)src->n_blocks);
    csgn_buf *b = nullptr;
    int rc = new_buf(n_blocks, src->L, 0, &b);
    if (rc != CSGN_OK) return rc;
    await_upload(src);
    cudaError_t e = launch_concat(src->d + first_block * src->L, n_blocks * src->L, nullptr, 0, b->d, g.stream);
    if (e != cudaSuccess) {
        csgn_buf_free(b);
        return cuda_fail(e, "slice kernel");
    }
    *out = b;
    return CSGN_OK;
}

int csgn_buf_download_range(const csgn_buf *buf, uint64_t first_block, uint64_t n_blocks, uint64_t *host_words) {
    NEED_INIT();
    if (!buf) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (first_block > buf->n_blocks || n_blocks > buf->n_blocks - first_block)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "block range [%llu,+%llu) outside %llu blocks",
                    (unsigned long long)first_block, (unsigned long long)n_blocks, (unsigned long long)buf->n_blocks);
    if (n_blocks == 0) return CSGN_OK;
    if (!host_words) return fail(CSGN_ERR_INVALID_ARGUMENT, "null host destination");
    await_upload(buf);
    CU(cudaMemcpyAsync(host_words, buf->d + first_block * buf->L, n_blocks * buf->L * sizeof(uint64_t),
                       cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    return CSGN_OK;
}

int csgn_buf_download(const csgn_buf *buf, uint64_t *host_words) {
    if (!buf) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    return csgn_buf_download_range(buf, 0, buf->n_blocks, host_words);
}

int csgn_buf_free(csgn_buf *buf) {
    if (!buf) return CSGN_OK;
    if (g.inited) {
        if (buf->ready) await_upload(buf);     // a never-consumed upload must land before its memory is recycled
        else order_after_last_use(buf);
    }
    if (g.inited && buf->owns && !(buf->recycle && give_upload_slot(buf->d, buf->cap_words))) dev_free(buf->d);
    delete buf;
    return CSGN_OK;
}

uint64_t csgn_buf_blocks(const csgn_buf *buf) { return buf ? buf->n_blocks : 0; }
uint32_t csgn_buf_words_per_block(const csgn_buf *buf) { return buf ? buf->L : 0; }
void *csgn_buf_device_ptr(const csgn_buf *buf) {
    if (buf && g.inited) await_upload(buf);   // whoever reads the pointer is ordered after the work stream
    return buf ? buf->d : nullptr;
}

// ---------------------------------------------------------------------------
// K1 multiply
// ---------------------------------------------------------------------------
int csgn_mul_into(const csgn_buf *a, const csgn_buf *b, csgn_buf *out) {
    NEED_INIT();
    if (!a || !b || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (a->L != b->L || a->L != out->L)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "words per block differ (%u, %u, %u)", a->L, b->L, out->L);
    if (b->n_blocks && a->n_blocks > UINT64_MAX / b->n_blocks)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "product block count overflows");
    if (out->n_blocks != a->n_blocks * b->n_blocks)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "output holds %llu blocks, product has %llu",
                    (unsigned long long)out->n_blocks, (unsigned long long)(a->n_blocks * b->n_blocks));
    if (out->d == a->d || out->d == b->d) return fail(CSGN_ERR_INVALID_ARGUMENT, "output aliases an operand");
    await_upload(a);
    await_upload(b);
    await_upload(out);
    cudaError_t e = launch_mul(a->d, a->n_blocks, b->d, b->n_blocks, a->L, out->d, g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "multiply kernel");
    return CSGN_OK;
}

int csgn_mul(const csgn_buf *a, const csgn_buf *b, csgn_buf **out) {
    NEED_INIT();
    if (!a || !b || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (a->L != b->L) return fail(CSGN_ERR_SHAPE_MISMATCH, "words per block differ (%u vs %u)", a->L, b->L);
    if (b->n_blocks && a->n_blocks > UINT64_MAX / b->n_blocks)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "product block count overflows");
    csgn_buf *c = nullptr;
    int rc = new_buf(a->n_blocks * b->n_blocks, a->L, 0, &c);
    if (rc != CSGN_OK) return rc;
    rc = csgn_mul_into(a, b, c);
    if (rc != CSGN_OK) {
        csgn_buf_free(c);
        return rc;
    }
    *out = c;
    return CSGN_OK;
}

// ---------------------------------------------------------------------------
// K2 add
// ---------------------------------------------------------------------------
int csgn_concat(const csgn_buf *a, const csgn_buf *b, csgn_buf **out) {
    NEED_INIT();
    if (!a || !b || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (a->L != b->L) return fail(CSGN_ERR_SHAPE_MISMATCH, "words per block differ (%u vs %u)", a->L, b->L);
    csgn_buf *c = nullptr;
    int rc = new_buf(a->n_blocks + b->n_blocks, a->L, 0, &c);
    if (rc != CSGN_OK) return rc;
    await_upload(a);
    await_upload(b);
    cudaError_t e = launch_concat(a->d, a->n_blocks * a->L, b->d, b->n_blocks * b->L, c->d, g.stream);
    if (e != cudaSuccess) {
        csgn_buf_free(c);
        return cuda_fail(e, "concat kernel");
    }
    *out = c;
    return CSGN_OK;
}

int csgn_append(csgn_buf *a, const csgn_buf *b) {
    NEED_INIT();
    if (!a || !b) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (!a->owns) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot grow a non-owning view");
    if (a->L != b->L) return fail(CSGN_ERR_SHAPE_MISMATCH, "words per block differ (%u vs %u)", a->L, b->L);
    const uint64_t na = a->n_blocks * a->L, nb = b->n_blocks * b->L;
    if (nb == 0) return CSGN_OK;
    await_upload(a);
    await_upload(b);
    if (na + nb <= a->cap_words) {
        // b == a is fine: source [0,na) and destination [na,2na) do not overlap
        cudaError_t e = launch_concat(a->d, na, b->d, nb, a->d, g.stream);
        if (e != cudaSuccess) return cuda_fail(e, "append kernel");
    } else {
        const uint64_t cap = std::max<uint64_t>(na + nb, 2 * a->cap_words);
        uint64_t *nd = nullptr;
        int rc = dev_alloc(cap, &nd);
        if (rc != CSGN_OK && cap > na + nb) rc = dev_alloc(na + nb, &nd);  // no room to double
        if (rc != CSGN_OK) return rc;
        cudaError_t e = launch_concat(a->d, na, b->d, nb, nd, g.stream);
        if (e != cudaSuccess) {
            dev_free(nd);
            return cuda_fail(e, "append kernel");
        }
        dev_free(a->d);  // stream-ordered: the copy above still reads it safely
        a->d = nd;
        a->cap_words = cap;
        a->recycle = false;
    }
    a->n_blocks += b->n_blocks;
    return CSGN_OK;
}

// ---------------------------------------------------------------------------
// K3 decrypt
// ---------------------------------------------------------------------------
int csgn_key_create(uint64_t N, const uint64_t *positions, uint32_t D, csgn_key **out) {
    NEED_INIT();
    if (!out || N == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad key arguments");
    if (D && !positions) return fail(CSGN_ERR_INVALID_ARGUMENT, "null positions");
    const uint32_t L = csgn_words_per_block(N);
    std::vector<uint64_t> mask(L, 0);
    for (uint32_t i = 0; i < D; ++i) {
        if (positions[i] >= N)
            return fail(CSGN_ERR_INVALID_ARGUMENT, "secret position %llu outside [0,%llu)",
                        (unsigned long long)positions[i], (unsigned long long)N);
        mask[positions[i] >> 6] |= 1ull << (63u - (positions[i] & 63u));
    }
    csgn_key *k = new csgn_key;
    k->N = N;
    k->L = L;
    k->D = D;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&k->d_mask), (size_t)L * sizeof(uint64_t));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(k->d_mask, mask.data(), (size_t)L * sizeof(uint64_t), cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess && D) e = cudaMalloc(reinterpret_cast<void **>(&k->d_positions), (size_t)D * sizeof(uint64_t));
    if (e == cudaSuccess && D)
        e = cudaMemcpyAsync(k->d_positions, positions, (size_t)D * sizeof(uint64_t), cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);  // `mask` dies with this frame
    if (e == cudaSuccess) k->h_mask = mask;
    std::fill(mask.begin(), mask.end(), 0);
    if (e != cudaSuccess) {
        if (k->d_mask) cudaFree(k->d_mask);
        if (k->d_positions) cudaFree(k->d_positions);
        delete k;
        return cuda_fail(e, "key upload");
    }
    *out = k;
    return CSGN_OK;
}

int csgn_key_free(csgn_key *key) {
    if (!key) return CSGN_OK;
    if (g.inited && key->d_mask) {
        cudaMemsetAsync(key->d_mask, 0, (size_t)key->L * sizeof(uint64_t), g.stream);  // zeroise, as the reference's dtor does
        if (key->d_positions) cudaMemsetAsync(key->d_positions, 0, (size_t)key->D * sizeof(uint64_t), g.stream);
        cudaStreamSynchronize(g.stream);
        cudaFree(key->d_mask);
        if (key->d_positions) cudaFree(key->d_positions);
    }
    std::fill(key->h_mask.begin(), key->h_mask.end(), 0);
    delete key;
    return CSGN_OK;
}

int csgn_decrypt_count_async(const csgn_buf *c, const csgn_key *key, uint64_t *device_count) {
    NEED_INIT();
    if (!c || !key || !device_count) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (c->L != key->L)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "ciphertext has %u words per block, key expects %u", c->L, key->L);
    await_upload(c);
    cudaError_t e = launch_decrypt_count(c->d, c->n_blocks, c->L, key->d_mask,
                                         key->h_mask.empty() ? nullptr : key->h_mask.data(), next_fold_scratch(), device_count,
                                         g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "decrypt kernel");
    return CSGN_OK;
}

int csgn_decrypt_count(const csgn_buf *c, const csgn_key *key, uint64_t *count) {
    NEED_INIT();
    if (!count) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output");
    int rc = csgn_decrypt_count_async(c, key, g.d_scratch + 2);
    if (rc != CSGN_OK) return rc;
    CU(cudaMemcpyAsync(g.h_result, g.d_scratch + 2, sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    *count = g.h_result[0];
    return CSGN_OK;
}

int csgn_decrypt(const csgn_buf *c, const csgn_key *key, uint8_t *bit) {
    if (!bit) return fail(CSGN_ERR_INVALID_ARGUMENT, "null output");
    uint64_t count = 0;
    int rc = csgn_decrypt_count(c, key, &count);
    if (rc != CSGN_OK) return rc;
    *bit = (uint8_t)(count & 1u);
    return CSGN_OK;
}

int csgn_decrypt_product(const csgn_buf *const *factors, uint32_t n_factors, const csgn_key *key, uint8_t *bit,
                         uint64_t *count) {
    NEED_INIT();
    if (!factors || !key || !bit || n_factors == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad product arguments");
    for (uint32_t i = 0; i < n_factors; ++i) {
        if (!factors[i]) return fail(CSGN_ERR_INVALID_ARGUMENT, "null factor %u", i);
        if (factors[i]->L != key->L)
            return fail(CSGN_ERR_SHAPE_MISMATCH, "factor %u has %u words per block, key expects %u", i, factors[i]->L,
                        key->L);
    }
    uint64_t *d_counts = nullptr;
    int rc = dev_alloc(n_factors, &d_counts);
    if (rc != CSGN_OK) return rc;
    for (uint32_t i = 0; i < n_factors; ++i) {
        rc = csgn_decrypt_count_async(factors[i], key, d_counts + i);
        if (rc != CSGN_OK) {
            dev_free(d_counts);
            return rc;
        }
    }
    std::vector<uint64_t> h(n_factors);
    cudaError_t e = cudaMemcpyAsync(h.data(), d_counts, n_factors * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    dev_free(d_counts);
    if (e != cudaSuccess) return cuda_fail(e, "product decrypt readback");
    uint64_t prod = 1;
    uint8_t parity = 1;
    for (uint32_t i = 0; i < n_factors; ++i) {
        parity &= (uint8_t)(h[i] & 1u);
        prod = (h[i] != 0 && prod > UINT64_MAX / h[i]) ? UINT64_MAX : prod * h[i];
    }
    *bit = parity;
    if (count) *count = prod;
    return CSGN_OK;
}

int csgn_decrypt_positions(const csgn_buf *c, uint64_t N, const uint64_t *positions, uint32_t D, uint8_t *bit) {
    csgn_key *k = nullptr;
    int rc = csgn_key_create(N, positions, D, &k);
    if (rc != CSGN_OK) return rc;
    rc = csgn_decrypt(c, k, bit);
    csgn_key_free(k);
    return rc;
}


// ---------------------------------------------------------------------------
// batches of independent items
// ---------------------------------------------------------------------------
int csgn_mul_into_batch(const csgn_buf *const *a, const csgn_buf *const *b, uint32_t n, csgn_buf *const *out) {
    NEED_INIT();
    if (n && (!a || !b || !out)) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    LaneScope lanes(n);
    for (uint32_t i = 0; i < n; ++i) {
        lanes.enter(i);
        int rc = csgn_mul_into(a[i], b[i], out[i]);
        if (rc != CSGN_OK) return rc;        // ~LaneScope joins what was enqueued
    }
    return CSGN_OK;
}

int csgn_mul_batch(const csgn_buf *const *a, const csgn_buf *const *b, uint32_t n, csgn_buf **out) {
    NEED_INIT();
    if (n && (!a || !b || !out)) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    for (uint32_t i = 0; i < n; ++i) out[i] = nullptr;
    LaneScope lanes(n);
    for (uint32_t i = 0; i < n; ++i) {
        lanes.enter(i);
        int rc = csgn_mul(a[i], b[i], &out[i]);
        if (rc != CSGN_OK) {
            lanes.join();
            for (uint32_t k = 0; k < i; ++k) {
                csgn_buf_free(out[k]);
                out[k] = nullptr;
            }
            return rc;
        }
    }
    return CSGN_OK;
}

int csgn_decrypt_count_batch_async(const csgn_buf *const *c, uint32_t n, const csgn_key *key, uint64_t *device_counts) {
    NEED_INIT();
    if (n && (!c || !device_counts)) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    LaneScope lanes(n);
    for (uint32_t i = 0; i < n; ++i) {
        lanes.enter(i);
        int rc = csgn_decrypt_count_async(c[i], key, device_counts + i);
        if (rc != CSGN_OK) return rc;
    }
    return CSGN_OK;
}

int csgn_decrypt_batch(const csgn_buf *const *c, uint32_t n, const csgn_key *key, uint8_t *bits, uint64_t *counts) {
    NEED_INIT();
    if (n == 0) return CSGN_OK;
    if (!bits && !counts) return fail(CSGN_ERR_INVALID_ARGUMENT, "no output");
    uint64_t *d = nullptr;
    int rc = dev_alloc(n, &d);
    if (rc != CSGN_OK) return rc;
    rc = csgn_decrypt_count_batch_async(c, n, key, d);
    std::vector<uint64_t> h(n);
    cudaError_t e = cudaSuccess;
    if (rc == CSGN_OK) {
        e = cudaMemcpyAsync(h.data(), d, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    }
    dev_free(d);
    if (rc != CSGN_OK) return rc;
    if (e != cudaSuccess) return cuda_fail(e, "batch decrypt readback");
    for (uint32_t i = 0; i < n; ++i) {
        if (bits) bits[i] = (uint8_t)(h[i] & 1u);
        if (counts) counts[i] = h[i];
    }
    return CSGN_OK;
}


// ---------------------------------------------------------------------------
// batched encryption
// ---------------------------------------------------------------------------
int csgn_encrypt_batch(const csgn_key *key, const uint8_t *bits, uint64_t n, uint64_t first_block, uint64_t seed,
                       csgn_buf **out) {
    NEED_INIT();
    if (!key || !out || (n && !bits)) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    if (key->D == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "the key has no secret positions");
    csgn_buf *b = nullptr;
    int rc = new_buf(n, key->L, 0, &b);
    if (rc != CSGN_OK) return rc;
    if (n) {
        uint8_t *d_bits = nullptr;
        cudaError_t e = cudaMallocAsync(reinterpret_cast<void **>(&d_bits), n, g.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_bits, bits, n, cudaMemcpyHostToDevice, g.stream);
        const uint64_t rem = key->N % 64;
        const uint64_t pad = rem ? ~0ull << (64 - rem) : ~0ull;
        if (e == cudaSuccess)
            e = launch_encrypt_batch(d_bits, n, first_block, key->L, pad, key->d_mask, key->d_positions, key->D, seed,
                                     b->d, g.stream);
        if (d_bits) cudaFreeAsync(d_bits, g.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);   // `bits` may be a temporary of the caller
        if (e != cudaSuccess) {
            csgn_buf_free(b);
            return cuda_fail(e, "batched encryption");
        }
    }
    *out = b;
    return CSGN_OK;
}

// ---------------------------------------------------------------------------
// K4 permute
// ---------------------------------------------------------------------------
int csgn_perm_create(uint64_t N, const uint64_t *perm, csgn_perm **out) {
    NEED_INIT();
    if (!out || !perm || N == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad permutation arguments");
    if (N >= (1ull << 32)) return fail(CSGN_ERR_INVALID_ARGUMENT, "N too large for the source map");
    std::vector<uint32_t> map(N);
    std::vector<uint8_t> seen(N, 0);
    for (uint64_t i = 0; i < N; ++i) {
        const uint64_t p = perm[i];
        if (p >= N || seen[p])
            return fail(CSGN_ERR_INVALID_ARGUMENT, "not a permutation of [0,%llu): entry %llu = %llu",
                        (unsigned long long)N, (unsigned long long)i, (unsigned long long)p);
        seen[p] = 1;
        map[i] = (uint32_t)(((p >> 6) << 6) | (63u - (p & 63u)));
    }
    csgn_perm *h = new csgn_perm;
    h->N = N;
    h->L = csgn_words_per_block(N);
    // Bit-sliced form: a block is W = 2L 32-bit words; bit j of word c holds position
    // 64*(c>>1) + (c odd ? 31-j : 63-j).  Entry j*W + c' names the source slice of output
    // (c', j) as a shared-memory BYTE offset 4*(stride*c + j_src), or the zero slot 4*stride*W for pad bits.
    std::vector<uint32_t> slices;
    if (permute_sliced_supported(h->L)) {
        const uint32_t W = 2 * h->L, stride = kPermSliceStride;
        slices.assign((size_t)32 * W, 4u * stride * W);   // byte offsets; default = the zero slot
        for (uint32_t c = 0; c < W; ++c)
            for (uint32_t j = 0; j < 32; ++j) {
                const uint64_t i = 64ull * (c >> 1) + ((c & 1u) ? 31u - j : 63u - j);
                if (i >= N) continue;
                const uint64_t p = perm[i];
                const uint32_t sc = 2u * (uint32_t)(p >> 6) + (((p & 63u) < 32u) ? 1u : 0u);
                slices[(size_t)j * W + c] = 4u * (stride * sc + (31u - (uint32_t)(p & 31u)));
            }
    }
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&h->d_map), (size_t)N * sizeof(uint32_t));
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(h->d_map, map.data(), (size_t)N * sizeof(uint32_t), cudaMemcpyHostToDevice, g.stream);
    if (e == cudaSuccess && !slices.empty()) {
        e = cudaMalloc(reinterpret_cast<void **>(&h->d_slice_map), slices.size() * sizeof(uint32_t));
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(h->d_slice_map, slices.data(), slices.size() * sizeof(uint32_t), cudaMemcpyHostToDevice,
                                g.stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess) {
        if (h->d_map) cudaFree(h->d_map);
        if (h->d_slice_map) cudaFree(h->d_slice_map);
        delete h;
        return cuda_fail(e, "permutation upload");
    }
    *out = h;
    return CSGN_OK;
}

int csgn_perm_free(csgn_perm *perm) {
    if (!perm) return CSGN_OK;
    if (g.inited && perm->d_map) {
        cudaStreamSynchronize(g.stream);
        cudaFree(perm->d_map);
        if (perm->d_slice_map) cudaFree(perm->d_slice_map);
    }
    delete perm;
    return CSGN_OK;
}

int csgn_permute_into(const csgn_buf *c, const csgn_perm *perm, csgn_buf *out) {
    NEED_INIT();
    if (!c || !perm || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (c->L != perm->L || out->L != perm->L)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "words per block: ciphertext %u, output %u, permutation %u", c->L,
                    out->L, perm->L);
    if (out->n_blocks > c->n_blocks)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "output holds more blocks than the input");
    if (out->d == c->d) return fail(CSGN_ERR_INVALID_ARGUMENT, "permute cannot run in place");
    await_upload(c);
    await_upload(out);
    cudaError_t e = launch_permute(c->d, out->n_blocks, c->L, (uint32_t)perm->N, perm->d_map, perm->d_slice_map, out->d,
                                   g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "permute kernel");
    return CSGN_OK;
}

int csgn_permute(const csgn_buf *c, const csgn_perm *perm, int strict_ref_truncate, csgn_buf **out) {
    NEED_INIT();
    if (!c || !perm || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    if (c->n_blocks == 0) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot permute an empty ciphertext");
    csgn_buf *r = nullptr;
    int rc = new_buf(strict_ref_truncate ? 1 : c->n_blocks, c->L, 0, &r);
    if (rc != CSGN_OK) return rc;
    rc = csgn_permute_into(c, perm, r);
    if (rc != CSGN_OK) {
        csgn_buf_free(r);
        return rc;
    }
    *out = r;
    return CSGN_OK;
}

// ---------------------------------------------------------------------------
// checksum, sharding
// ---------------------------------------------------------------------------
int csgn_buf_checksum(const csgn_buf *buf, uint64_t *xor_out, uint64_t *sum_out, uint64_t *wsum_out) {
    NEED_INIT();
    if (!buf) return fail(CSGN_ERR_INVALID_ARGUMENT, "null handle");
    uint64_t *acc = g.d_scratch + 4;
    await_upload(buf);
    CU(cudaMemsetAsync(acc, 0, 3 * sizeof(uint64_t), g.stream));
    cudaError_t e = launch_checksum(buf->d, buf->n_blocks * buf->L, acc, g.stream);
    if (e != cudaSuccess) return cuda_fail(e, "checksum kernel");
    CU(cudaMemcpyAsync(g.h_result + 4, acc, 3 * sizeof(uint64_t), cudaMemcpyDeviceToHost, g.stream));
    CU(cudaStreamSynchronize(g.stream));
    if (xor_out) *xor_out = g.h_result[4];
    if (sum_out) *sum_out = g.h_result[5];
    if (wsum_out) *wsum_out = g.h_result[6];
    return CSGN_OK;
}

int csgn_shard_range(uint64_t n_blocks, int rank, int world, uint64_t *first, uint64_t *count) {
    if (world < 1 || rank < 0 || rank >= world || !first || !count)
        return fail(CSGN_ERR_INVALID_ARGUMENT, "bad shard arguments (rank %d of %d)", rank, world);
    const uint64_t base = n_blocks / (uint64_t)world, extra = n_blocks % (uint64_t)world;
    const uint64_t r = (uint64_t)rank;
    *first = r * base + std::min<uint64_t>(r, extra);
    *count = base + (r < extra ? 1 : 0);
    return CSGN_OK;
}

}  // extern "C"
