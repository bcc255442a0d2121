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

namespace {
int save_impl(const csgn_buf *buf, uint64_t N, uint64_t D, const char *path, uint64_t r0, uint64_t r1);
int load_impl(const char *path, uint64_t *N, uint64_t *D, uint64_t *r0, uint64_t *r1, csgn_buf **out);
std::string shard_path(const char *prefix, int rank, int world) {
    char tail[64];
    snprintf(tail, sizeof tail, ".shard%dof%d", rank, world);
    return std::string(prefix) + tail;
}
}  // namespace

int csgn_buf_save(const csgn_buf *buf, uint64_t N, uint64_t D, const char *path) { return save_impl(buf, N, D, path, 0, 0); }

int csgn_buf_load(const char *path, uint64_t *N, uint64_t *D, csgn_buf **out) {
    uint64_t r0 = 0, r1 = 0;
    int rc = load_impl(path, N, D, &r0, &r1, out);
    if (rc == CSGN_OK && r0 != 0) {
        csgn_buf_free(*out);
        *out = nullptr;
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s is one shard of a sharded ciphertext (rank %u of %u): use csgn_buf_load_shard",
                    path, (unsigned)(r0 & 0xffffffffu) - 1u, (unsigned)(r0 >> 32));
    }
    return rc;
}

// One file per rank: `<prefix>.shard<rank>of<world>`, the ordinary ciphertext file of the rank's local blocks with
// (rank + 1 | world << 32) and the first global block in the header's reserved words.
int csgn_buf_save_shard(const csgn_buf *buf, uint64_t N, uint64_t D, const char *prefix, int rank, int world,
                        uint64_t first_block) {
    if (!prefix || world < 1 || rank < 0 || rank >= world) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad shard arguments");
    return save_impl(buf, N, D, shard_path(prefix, rank, world).c_str(), (uint64_t)(rank + 1) | ((uint64_t)world << 32), first_block);
}

int csgn_buf_load_shard(const char *prefix, int rank, int world, uint64_t *N, uint64_t *D, uint64_t *first_block,
                        csgn_buf **out) {
    if (!prefix || world < 1 || rank < 0 || rank >= world) return fail(CSGN_ERR_INVALID_ARGUMENT, "bad shard arguments");
    const std::string path = shard_path(prefix, rank, world);
    uint64_t r0 = 0, r1 = 0;
    int rc = load_impl(path.c_str(), N, D, &r0, &r1, out);
    if (rc != CSGN_OK) return rc;
    if (r0 != ((uint64_t)(rank + 1) | ((uint64_t)world << 32))) {
        csgn_buf_free(*out);
        *out = nullptr;
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s was not written as shard %d of %d", path.c_str(), rank, world);
    }
    if (first_block) *first_block = r1;
    return CSGN_OK;
}

namespace {
int save_impl(const csgn_buf *buf, uint64_t N, uint64_t D, const char *path, uint64_t r0, uint64_t r1) {
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
    h.reserved[0] = r0; h.reserved[1] = r1;
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

int load_impl(const char *path, uint64_t *N, uint64_t *D, uint64_t *r0, uint64_t *r1, csgn_buf **out) {
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
    *r0 = h.reserved[0];
    *r1 = h.reserved[1];
    *out = b;
    return CSGN_OK;
}
}  // namespace

// ---------------------------------------------------------------------------
// SecretKey / Permutation files (host only: no device is touched, csgn_init is not required)
// 64-byte header {magic, N, D, count, xor of the entries, 0...} + count uint64 entries.
// ---------------------------------------------------------------------------
namespace {
const char kKeyMagic[8] = {'C', 'S', 'G', 'N', 'S', 'K', '0', '1'};
const char kPermMagic[8] = {'C', 'S', 'G', 'N', 'P', 'M', '0', '1'};

int words_save(const char *path, const char *magic, uint64_t N, uint64_t D, const uint64_t *w, uint64_t n) {
    if (!path || (n && !w)) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    FILE *f = fopen(path, "wb");
    if (!f) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot open %s for writing", path);
    FileHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, magic, 8);
    h.N = N; h.D = D; h.L = 0; h.n_blocks = n; h.xor_words = xor_fold(w, n);
    bool ok = fwrite(&h, sizeof h, 1, f) == 1 && fwrite(w, 8, n, f) == n;
    ok = (fclose(f) == 0) && ok;
    return ok ? CSGN_OK : fail(CSGN_ERR_INVALID_ARGUMENT, "short write to %s", path);
}

int words_load(const char *path, const char *magic, const char *what, uint64_t *N, uint64_t *D, uint64_t *w, uint64_t cap,
               uint64_t *n) {
    if (!path || !n) return fail(CSGN_ERR_INVALID_ARGUMENT, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(CSGN_ERR_INVALID_ARGUMENT, "cannot open %s", path);
    FileHeader h;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, magic, 8) != 0) {
        fclose(f);
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s is not a CSGN %s file", path, what);
    }
    *n = h.n_blocks;
    if (N) *N = h.N;
    if (D) *D = h.D;
    if (!w) {                      // size query
        fclose(f);
        return CSGN_OK;
    }
    if (h.n_blocks > cap) {
        fclose(f);
        return fail(CSGN_ERR_INVALID_ARGUMENT, "%s holds %llu entries, room for %llu", path, (unsigned long long)h.n_blocks,
                    (unsigned long long)cap);
    }
    const bool ok = fread(w, 8, h.n_blocks, f) == h.n_blocks;
    fclose(f);
    if (!ok) return fail(CSGN_ERR_INVALID_ARGUMENT, "%s: truncated file", path);
    if (xor_fold(w, h.n_blocks) != h.xor_words) return fail(CSGN_ERR_INVALID_ARGUMENT, "%s: checksum mismatch (file corrupted)", path);
    return CSGN_OK;
}
}  // namespace

int csgn_key_positions_save(const char *path, uint64_t N, uint64_t D, const uint64_t *positions, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i)
        if (positions && positions[i] >= N) return fail(CSGN_ERR_INVALID_ARGUMENT, "secret position %llu outside [0, N)", (unsigned long long)positions[i]);
    return words_save(path, kKeyMagic, N, D, positions, n);
}

int csgn_key_positions_load(const char *path, uint64_t *N, uint64_t *D, uint64_t *positions, uint64_t capacity, uint64_t *n) {
    uint64_t nn = 0, NN = 0;
    int rc = words_load(path, kKeyMagic, "secret key", &NN, D, positions, capacity, &nn);
    if (rc != CSGN_OK) return rc;
    if (positions)
        for (uint64_t i = 0; i < nn; ++i)
            if (positions[i] >= NN) return fail(CSGN_ERR_INVALID_ARGUMENT, "%s: secret position outside [0, N)", path);
    if (N) *N = NN;
    if (n) *n = nn;
    return CSGN_OK;
}

int csgn_perm_entries_save(const char *path, const uint64_t *perm, uint64_t n) { return words_save(path, kPermMagic, n, 0, perm, n); }

int csgn_perm_entries_load(const char *path, uint64_t *perm, uint64_t capacity, uint64_t *n) {
    uint64_t nn = 0;
    int rc = words_load(path, kPermMagic, "permutation", nullptr, nullptr, perm, capacity, &nn);
    if (rc != CSGN_OK) return rc;
    if (perm) {
        std::vector<uint8_t> seen(nn, 0);
        for (uint64_t i = 0; i < nn; ++i) {
            if (perm[i] >= nn || seen[perm[i]]) return fail(CSGN_ERR_INVALID_ARGUMENT, "%s: not a permutation of [0,%llu)", path, (unsigned long long)nn);
            seen[perm[i]] = 1;
        }
    }
    if (n) *n = nn;
    return CSGN_OK;
}

}  // extern "C"
