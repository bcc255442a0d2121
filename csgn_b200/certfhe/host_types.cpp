// host_types.cpp -- the host-only value types of the certFHE API: Library, Helper,
// Context, Plaintext, Permutation, Timer.  Behaviour follows the reference file by
// file (citations inline); where the reference's random generation defines results
// (Permutation(size)), glibc rand() is consumed in exactly its order.
#include "certFHE.h"

#include <vector>
#include "engine_glue.h"

#include <ctime>

namespace certFHE {

// ---- Library ---------------------------------------------------------------------------
namespace {
bool g_strict_permute = false;
bool g_strict_permute_set = false;
bool g_lazy_products = false;
bool g_lazy_products_set = false;
bool g_fused_products = true;
bool g_fused_products_set = false;
bool g_auto_lanes = true;
bool g_auto_lanes_set = false;
bool g_rope_sums = true;
bool g_rope_sums_set = false;

bool env_flag(const char *name, bool dflt) {
    const char *e = getenv(name);
    if (!e || !*e) return dflt;
    return strcmp(e, "0") != 0;
}
}  // namespace

void Library::initializeLibrary() {
    // reference src/Helpers.cpp:8-12: the PRNG is seeded with the local time
    srand(time(NULL));
    glue::ensure_engine();
    glue::check(csgn_set_auto_lanes(getAutoLanes() ? 1 : 0), "csgn_set_auto_lanes");
    // one process per GPU under a launcher: join the job's peers
    const char *ws = getenv("WORLD_SIZE"), *rk = getenv("RANK"), *dir = getenv("CSGN_RENDEZVOUS_DIR");
    if (ws && rk && dir && atoi(ws) >= 1 && !glue::comm()) {
        const char *tag = getenv("CSGN_JOB_TAG");
        if (!tag || !*tag) tag = getenv("MASTER_PORT");
        connectPeers(atoi(rk), atoi(ws), dir, tag && *tag ? tag : "job");
    }
}

namespace {
csgn_comm *g_comm = nullptr;
int g_rank = 0, g_world = 1;
}  // namespace

namespace glue {
csgn_comm *comm() { return g_comm; }
}  // namespace glue

void Library::connectPeers(int rank, int world, const std::string &rendezvous_dir, const std::string &job_tag) {
    glue::ensure_engine();
    if (g_comm) throw Error("Library::connectPeers: already connected");
    unsigned char handle[CSGN_IPC_HANDLE_BYTES];
    csgn_comm *c = nullptr;
    glue::check(csgn_comm_create(rank, world, &c, handle), "csgn_comm_create");
    const char *t = getenv("CSGN_RENDEZVOUS_TIMEOUT_MS");
    const int rc = csgn_comm_connect_dir(c, handle, rendezvous_dir.c_str(), job_tag.c_str(), t && *t ? atoi(t) : 60000);
    if (rc != CSGN_OK) {
        const std::string why = csgn_last_error();
        csgn_comm_free(c);
        throw Error("Library::connectPeers: " + why);
    }
    g_comm = c;
    g_rank = rank;
    g_world = world;
    atexit([] {                      // closes the peer mappings and removes this rank's handle file
        if (g_comm) csgn_comm_free(g_comm);
        g_comm = nullptr;
    });
}

int Library::getRank() { return g_rank; }
int Library::getWorldSize() { return g_world; }

void Library::setStrictReferencePermutation(bool strict) {
    g_strict_permute = strict;
    g_strict_permute_set = true;
}

bool Library::getStrictReferencePermutation() {
    if (!g_strict_permute_set) {
        const char *e = getenv("CSGN_STRICT_REF_PERMUTE");
        g_strict_permute = e && *e && strcmp(e, "0") != 0;
        g_strict_permute_set = true;
    }
    return g_strict_permute;
}

void Library::setLazyProducts(bool lazy) {
    g_lazy_products = lazy;
    g_lazy_products_set = true;
}

bool Library::getLazyProducts() {
    if (!g_lazy_products_set) {
        const char *e = getenv("CSGN_LAZY_PRODUCTS");
        g_lazy_products = e && *e && strcmp(e, "0") != 0;
        g_lazy_products_set = true;
    }
    return g_lazy_products;
}

void Library::setFusedProducts(bool fused) {
    g_fused_products = fused;
    g_fused_products_set = true;
}

bool Library::getFusedProducts() {
    if (!g_fused_products_set) {
        g_fused_products = env_flag("CSGN_FUSED_PRODUCTS", true);
        g_fused_products_set = true;
    }
    return g_fused_products;
}

void Library::setAutoLanes(bool on) {
    g_auto_lanes = on;
    g_auto_lanes_set = true;
    if (csgn_is_initialized()) glue::check(csgn_set_auto_lanes(on ? 1 : 0), "csgn_set_auto_lanes");
}

bool Library::getAutoLanes() {
    if (!g_auto_lanes_set) {
        g_auto_lanes = env_flag("CSGN_AUTO_LANES", true);
        g_auto_lanes_set = true;
    }
    return g_auto_lanes;
}

void Library::setRopeSums(bool on) {
    g_rope_sums = on;
    g_rope_sums_set = true;
}

bool Library::getRopeSums() {
    if (!g_rope_sums_set) {
        g_rope_sums = env_flag("CSGN_ROPE_SUMS", true);
        g_rope_sums_set = true;
    }
    return g_rope_sums;
}

void Library::synchronize() {
    glue::ensure_engine();
    glue::check(csgn_sync(), "csgn_sync");
}

// ---- Helper ----------------------------------------------------------------------------
bool Helper::exists(const uint64_t *v, const uint64_t len, const uint64_t value) {
    // reference src/Helpers.cpp:18-26: linear scan
    for (uint64_t i = 0; i < len; ++i)
        if (v[i] == value) return true;
    return false;
}

void Helper::deletePointer(void *pointer, bool isArray) {
    // reference src/Helpers.cpp:28-35 frees through a void* (ill-formed delete); every
    // array this library hands out is uint64_t[], so release it as that.
    if (!pointer) return;
    if (isArray) delete[] static_cast<uint64_t *>(pointer);
    else delete static_cast<uint64_t *>(pointer);
}

// ---- Context ---------------------------------------------------------------------------
static uint64_t words_for(uint64_t n) { return n / 64 + (n % 64 ? 1 : 0); }

Context::Context(const uint64_t pN, const uint64_t pD) : N(pN), D(pD) {
    // reference src/Context.cpp:20-29
    S = D ? N / (2 * D) : 0;
    defaultLen = words_for(N);
}

Context::Context(const Context &c) : N(c.N), D(c.D), S(c.S), defaultLen(c.defaultLen) {}

Context::~Context() {}

Context &Context::operator=(const Context &c) {
    N = c.N;
    D = c.D;
    S = c.S;
    defaultLen = c.defaultLen;
    return *this;
}

ostream &operator<<(ostream &out, const Context &c) {
    // reference src/Context.cpp:40-47
    out << "N= " << c.getN() << endl << "D= " << c.getD() << endl << "S= " << c.getS() << endl;
    return out;
}

uint64_t Context::getN() const { return N; }
uint64_t Context::getD() const { return D; }
uint64_t Context::getS() const { return S; }
uint64_t Context::getDefaultN() const { return defaultLen; }

void Context::setN(uint64_t n) {
    // the reference leaves defaultLen stale here (src/Context.cpp:81-85); we keep it in step
    N = n;
    S = D ? n / (2 * D) : 0;
    defaultLen = words_for(n);
}

void Context::setD(uint64_t d) {
    D = d;
    S = d ? N / (2 * d) : 0;
}

// ---- Plaintext -------------------------------------------------------------------------
Plaintext::Plaintext() : value(0) {}
Plaintext::Plaintext(const int v) : value((unsigned char)(v & 0x01)) {}  // src/Plaintext.cpp:30-33
Plaintext::~Plaintext() {}

void Plaintext::resolve() const {
    if (!pending) return;
    uint64_t count = 0;
    glue::check(csgn_result_wait(pending.get(), &count), "csgn_result_wait");
    value = (unsigned char)(count & 1u);      // the XOR over the blocks is the parity of the satisfied-block count
    pending.reset();
}

unsigned char Plaintext::getValue() const {
    resolve();
    return value;
}

void Plaintext::setValue(unsigned char v) {
    pending.reset();
    value = v & 0x01;
}

ostream &operator<<(ostream &out, const Plaintext &c) {
    // reference src/Plaintext.cpp:10-19: the digit, then a newline
    out << (char)('0' | c.getValue()) << endl;
    return out;
}

// ---- Permutation -----------------------------------------------------------------------
Permutation::Permutation() : permutation(nullptr), length(0), device_map(nullptr) {}

Permutation::Permutation(const uint64_t *perm, const uint64_t len) : Permutation() {
    length = len;
    permutation = new uint64_t[len ? len : 1];
    for (uint64_t i = 0; i < len; ++i) permutation[i] = perm[i];
}

Permutation::Permutation(const uint64_t size) : Permutation() {
    // reference src/Permutation.cpp:139-157: every slot starts at (uint64_t)-1; slot i
    // draws rand()%size until the value is not yet present anywhere in the array.
    // The reference's membership test is a linear scan (Helper::exists), O(n^2 log n) in all --
    // seconds at N=16383.  A taken-map answers the same question in O(1); the rand() draws, their
    // order and therefore the permutation are unchanged.
    length = size;
    permutation = new uint64_t[size ? size : 1];
    for (uint64_t i = 0; i < size; ++i) permutation[i] = (uint64_t)-1;
    std::vector<unsigned char> taken(size ? size : 1, 0);
    for (uint64_t i = 0; i < size; ++i) {
        uint64_t r = (uint64_t)rand() % size;
        while (taken[r]) r = (uint64_t)rand() % size;
        taken[r] = 1;
        permutation[i] = r;
    }
}

Permutation::Permutation(const Context &context) : Permutation(context.getN()) {}

void Permutation::save(const std::string &path) const {
    glue::check(csgn_perm_entries_save(path.c_str(), permutation, length), "csgn_perm_entries_save");
}

Permutation Permutation::load(const std::string &path) {
    uint64_t count = 0;
    glue::check(csgn_perm_entries_load(path.c_str(), nullptr, 0, &count), "csgn_perm_entries_load");
    std::vector<uint64_t> p(count ? count : 1);
    glue::check(csgn_perm_entries_load(path.c_str(), p.data(), count, &count), "csgn_perm_entries_load");
    return Permutation(p.data(), count);
}

Permutation::Permutation(const Permutation &p) : Permutation(p.permutation, p.length) {}

Permutation::~Permutation() {
    drop_device_map();
    delete[] permutation;
    permutation = nullptr;
    length = 0;
}

void Permutation::drop_device_map() const {
    if (device_map) {
        csgn_perm_free(device_map);
        device_map = nullptr;
    }
}

uint64_t Permutation::getLength() const { return length; }
uint64_t *Permutation::getPermutation() const { return permutation; }
void Permutation::setLength(uint64_t len) { length = len; }

void Permutation::setPermutation(uint64_t *perm, uint64_t len) {
    drop_device_map();
    uint64_t *fresh = new uint64_t[len ? len : 1];
    for (uint64_t i = 0; i < len; ++i) fresh[i] = perm[i];
    delete[] permutation;
    permutation = fresh;
    length = len;
}

Permutation &Permutation::operator=(const Permutation &p) {
    if (this != &p) setPermutation(p.permutation, p.length);
    return *this;
}

ostream &operator<<(ostream &out, const Permutation &p) {
    // reference src/Permutation.cpp:33-46: two parenthesised rows
    out << "(";
    for (uint64_t i = 0; i < p.getLength(); ++i) out << i << " ";
    out << ")" << endl << "(";
    for (uint64_t i = 0; i < p.getLength(); ++i) out << p.getPermutation()[i] << " ";
    out << ")" << endl;
    return out;
}

Permutation Permutation::getInverse() {
    // reference src/Permutation.cpp:8-27 searches j with permutation[j] == i for every i
    // (O(n^2)); scattering j to slot permutation[j] gives the same array in O(n).
    uint64_t *inv = new uint64_t[length ? length : 1];
    for (uint64_t j = 0; j < length; ++j)
        if (permutation[j] < length) inv[permutation[j]] = j;
    Permutation out(inv, length);
    delete[] inv;
    return out;
}

Permutation Permutation::operator+(const Permutation &b) const {
    // reference src/Permutation.cpp:63-78; length mismatch yields the empty permutation
    if (length != b.getLength()) return Permutation();
    Permutation out(permutation, length);
    for (uint64_t i = 0; i < length; ++i) out.permutation[i] = permutation[b.permutation[i]];
    return out;
}

Permutation &Permutation::operator+=(const Permutation &b) {
    // reference src/Permutation.cpp:80-96; length mismatch leaves *this untouched
    if (length != b.getLength()) return *this;
    uint64_t *p = new uint64_t[length ? length : 1];
    for (uint64_t i = 0; i < length; ++i) p[i] = permutation[b.permutation[i]];
    drop_device_map();
    delete[] permutation;
    permutation = p;
    return *this;
}

csgn_perm *Permutation::deviceMap(uint64_t N) const {
    if (length != N)
        throw Error("Permutation of length " + to_string(length) + " applied in a context with N = " + to_string(N));
    if (!device_map) {
        glue::ensure_engine();
        glue::check(csgn_perm_create(N, permutation, &device_map), "csgn_perm_create");
    }
    return device_map;
}

// ---- Timer -----------------------------------------------------------------------------
Timer::Timer(string pname)
    : name(pname), chronometer(0), start_fingerprint(std::chrono::high_resolution_clock::now()),
      stop_fingerprint(start_fingerprint) {}

Timer::~Timer() {}

void Timer::start() { start_fingerprint = std::chrono::high_resolution_clock::now(); }

double Timer::stop() {
    stop_fingerprint = std::chrono::high_resolution_clock::now();
    chronometer = stop_fingerprint - start_fingerprint;
    return chronometer.count() * 1000;
}

void Timer::reset() {
    stop_fingerprint = std::chrono::high_resolution_clock::now();
    start_fingerprint = stop_fingerprint;
}

void Timer::print() {
    // reference src/Timer.cpp:34-36: "name : X ms "
    cout << name << " : " << chronometer.count() * 1000 << " ms " << endl;
    fflush(stdout);
}

double Timer::stopAndPrint() {
    stop();
    print();
    return chronometer.count() * 1000;
}

double Timer::getValue() { return chronometer.count() * 1000; }

}  // namespace certFHE
