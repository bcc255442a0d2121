// capi_io.cu -- csgn_buf_save / csgn_buf_load: a ciphertext as a 64-byte header plus the reference's words verbatim,
// streamed between the device and the file through two pinned staging buffers (the reference has no
// serialisation, only size(), src/Ciphertext.cpp:91-101).
#include "capi_internal.cuh"

using namespace csgn;
using namespace csgn::detail;

extern "C" {

// ---------------------------------------------------------------------------
// serialisation
// ---------------------------------------------------------------------------
namespace {

struct FileHeader {
    char magic[8];
    uint64_t N, D, L, n_blocks, xor_words;
    uint64_t reserved[2];
};
static_assert(sizeof(FileHeader) == 64, "header is 64 bytes");
const char kMagic[8] = {'C', 'S', 'G', 'N', 'C', 'T', '0', '1'};
constexpr size_t kStageBytes = 32u << 20;   // two pinned staging buffers of 32 MiB

struct Staging {
    uint64_t *buf[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    ~Staging() {
        for (int i = 0; i < 2; ++i) {
            if (buf[i]) cudaFreeHost(buf[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
    }
    cudaError_t init() {
        for (int i = 0; i < 2; ++i) {
            cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&buf[i]), kStageBytes, cudaHostAllocDefault);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    }
};

uint64_t xor_fold(const uint64_t *w, size_t n) {
    uint64_t x = 0;
    for (size_t i = 0; i < n; ++i) x ^= w[i];
    return x;
}

}  // namespace

int csgn_buf_save(const csgn_buf *buf, uint64_t N, uint64_t D, const char *path) {
    NEED_INIT();
    if (!buf || !path) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    if (csgn_words_per_block(N) != buf->L)
        return fail(CSGN_ERR_SHAPE_MISMATCH, "N = %llu gives %u words per block, buffer has %u", (unsigned long long)N,
                    csgn_words_per_block(N), buf->L);
    {
        int rc = need_dense(buf);
        if (rc != CSGN_OK) return rc;
    }
    FILE *f = fopen(path, "wb");
    if (!f) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot open %s for writing", path);
    FileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.N = N; h.D = D; h.L = buf->L; h.n_blocks = buf->n_blocks;
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    Staging st;
    cudaError_t e = st.init();
    acquire_read(buf);
    const uint64_t total = buf->n_blocks * buf->L, per = kStageBytes / 8;
    uint64_t x = 0;
    // D2H of piece k+1 overlaps the fwrite of piece k
    uint64_t issued = 0, written = 0;
    int slot = 0;
    uint64_t len[2] = {0, 0};
    while (ok && e == cudaSuccess && written < total) {
        while (issued < total && issued - written < 2 * per) {
            const int s = (int)((issued / per) & 1);
            len[s] = std::min<uint64_t>(per, total - issued);
            e = cudaMemcpyAsync(st.buf[s], buf->d + issued, len[s] * 8, cudaMemcpyDeviceToHost, g.stream);
            if (e == cudaSuccess) e = cudaEventRecord(st.done[s], g.stream);
            if (e != cudaSuccess) break;
            issued += len[s];
        }
        if (e != cudaSuccess) break;
        e = cudaEventSynchronize(st.done[slot]);
        if (e != cudaSuccess) break;
        x ^= xor_fold(st.buf[slot], len[slot]);
        ok = fwrite(st.buf[slot], 8, len[slot], f) == len[slot];
        written += len[slot];
        slot ^= 1;
    }
    if (ok && e == cudaSuccess) {
        h.xor_words = x;
        ok = fseek(f, 0, SEEK_SET) == 0 && fwrite(&h, sizeof h, 1, f) == 1;
    }
    ok = (fclose(f) == 0) && ok;
    if (e != cudaSuccess) return cuda_fail(e, "save: device to host");
    if (!ok) return fail(CSGN_ERR_INVALID_ARGUMENT, "short write to %s", path);
    return CSGN_OK;
}

int csgn_buf_load(const char *path, uint64_t *N, uint64_t *D, csgn_buf **out) {
    NEED_INIT();
    if (!path || !out) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot open %s", path);
    FileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kMagic, 8) != 0) {
        fclose(f);
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s is not a CSGN ciphertext file", path);
    }
    if (h.L == 0 || h.L != csgn_words_per_block(h.N) || h.n_blocks > (UINT64_MAX / 8) / h.L) {
        fclose(f);
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s: inconsistent header (N=%llu L=%llu blocks=%llu)", path,
                    (unsigned long long)h.N, (unsigned long long)h.L, (unsigned long long)h.n_blocks);
    }
    csgn_buf *b = nullptr;
    int rc = new_buf(h.n_blocks, (uint32_t)h.L, 0, &b);
    if (rc != CSGN_OK) {
        fclose(f);
        return rc;
    }
    Staging st;
    cudaError_t e = st.init();
    const uint64_t total = h.n_blocks * h.L, per = kStageBytes / 8;
    uint64_t x = 0, done = 0;
    bool ok = true, used[2] = {false, false};
    int slot = 0;
    while (ok && e == cudaSuccess && done < total) {
        if (used[slot]) e = cudaEventSynchronize(st.done[slot]);   // the H2D that last read this buffer
        if (e != cudaSuccess) break;
        const uint64_t n = std::min<uint64_t>(per, total - done);
        ok = fread(st.buf[slot], 8, n, f) == n;
        if (!ok) break;
        x ^= xor_fold(st.buf[slot], n);
        e = cudaMemcpyAsync(b->d + done, st.buf[slot], n * 8, cudaMemcpyHostToDevice, g.stream);
        if (e == cudaSuccess) e = cudaEventRecord(st.done[slot], g.stream);
        used[slot] = true;
        done += n;
        slot ^= 1;
    }
    fclose(f);
    if (e == cudaSuccess) e = cudaStreamSynchronize(g.stream);
    if (e != cudaSuccess || !ok || x != h.xor_words) {
        csgn_buf_free(b);
        if (e != cudaSuccess) return cuda_fail(e, "load: host to device");
        return fail(CSGN_ERR_INVALID_ARGUMENT, ok ? "%s: checksum mismatch (file corrupted)" : "%s: truncated file", path);
    }
    if (N) *N = h.N;
    if (D) *D = h.D;
    *out = b;
    return CSGN_OK;
}

}  // extern "C"
