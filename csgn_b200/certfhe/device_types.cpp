// device_types.cpp -- Ciphertext and SecretKey of the certFHE API over the GPU engine.
//
// Ciphertext owns a csgn_buf (device words) instead of the reference's two host arrays
// (src/Ciphertext.h:17-21).  Each evaluation operator is ONE call into the C ABI:
//   operator+ / +=   -> csgn_concat / csgn_append     (src/Ciphertext.cpp:204-229, :249-281)
//   operator* / *=   -> csgn_mul                      (src/Ciphertext.cpp:231-247, :283-304)
//   applyPermutation -> csgn_permute                  (src/Ciphertext.cpp:7-89)
//   SecretKey::decrypt -> csgn_decrypt                (src/SecretKey.cpp:208-224)
// Encryption and key handling stay on the host and replay the reference's rand() order.
#include "certFHE.h"
#include "engine_glue.h"

#include <cstdlib>
#include <ctime>
#include <utility>
#include <vector>

namespace certFHE {

namespace glue {

void check(int status, const char *what) {
    if (status != CSGN_OK) throw Error(string(what) + ": " + csgn_last_error());
}

void ensure_engine() {
    if (csgn_is_initialized()) return;
    check(csgn_init(-1), "csgn_init");
    check(csgn_set_auto_lanes(Library::getAutoLanes() ? 1 : 0), "csgn_set_auto_lanes");
}

}  // namespace glue

namespace {

// canonical valid-bit count of word i of a ciphertext (src/SecretKey.cpp:171-173)
inline uint64_t canonical_bits(uint64_t i, uint64_t L, uint64_t rem) {
    return ((i % L) + 1 == L && rem) ? rem : 64;
}

void require_canonical_bitlen(const uint64_t *bitlen, uint64_t len, const Context &ctx) {
    // This runs in every Ciphertext(V, Bitlen, len, ctx), over as many words as V has: the first block is compared
    // with the canonical pattern and the rest of the array with itself one block earlier (the pattern has period L) --
    // one memcmp, which the C library runs on the widest vectors the host has (5.8 -> 3 us per 160 KB against the
    // hand-written loop, tools/ctor_probe.cpp).
    const uint64_t L = ctx.getDefaultN(), rem = ctx.getN() % 64;
    uint64_t diff = 0;
    const uint64_t head = len < L ? len : L;
    for (uint64_t k = 0; k < head; ++k) diff |= bitlen[k] ^ canonical_bits(k, L, rem);
    if (len > L && memcmp(bitlen + L, bitlen, (size_t)(len - L) * sizeof(uint64_t)) != 0) diff = 1;
    if (!diff) return;
    for (uint64_t i = 0; i < len; ++i)
        if (bitlen[i] != canonical_bits(i, L, rem))
            throw Error("Ciphertext: bitlen[" + to_string(i) + "] = " + to_string(bitlen[i]) +
                        " is not the canonical pattern for N = " + to_string(ctx.getN()) +
                        " (the reference would mis-index such an object, src/SecretKey.cpp:133)");
}

uint64_t *copy_words(const uint64_t *src, uint64_t n) {
    uint64_t *p = new uint64_t[n ? n : 1];
    if (n) memcpy(p, src, n * sizeof(uint64_t));
    return p;
}

}  // namespace

// =========================================================================================
// Ciphertext
// =========================================================================================
namespace {

std::shared_ptr<csgn_buf> adopt(csgn_buf *b) {
    return std::shared_ptr<csgn_buf>(b, [](csgn_buf *p) { csgn_buf_free(p); });
}

}  // namespace

Ciphertext::Ciphertext()
    : certFHEcontext(nullptr), host_v(nullptr), host_bitlen(nullptr), host_len(0), host_v_valid(false),
      staged(false), sharded(false) {}

Ciphertext::Ciphertext(const uint64_t *V, const uint64_t *Bitlen, const uint64_t len, const Context &context)
    : Ciphertext() {
    // reference src/Ciphertext.cpp:344-358: deep copy of the caller's arrays
    certFHEcontext = new Context(context);
    const uint64_t L = context.getDefaultN();
    if (L == 0 || len % L != 0)
        throw Error("Ciphertext: len = " + to_string(len) + " is not a multiple of the " + to_string(L) +
                    " words per block");
    if (Bitlen) require_canonical_bitlen(Bitlen, len, context);
    glue::ensure_engine();
    csgn_buf *b = nullptr;
    // the caller may free V right after the constructor: the words are copied into library-owned pinned staging
    // before this returns and travel to the device from there (no synchronisation)
    glue::check(csgn_buf_upload_copy(V, len / L, (uint32_t)L, &b), "csgn_buf_upload_copy");
    dev = adopt(b);
}

Ciphertext::Ciphertext(const Ciphertext &o) : Ciphertext() {
    // The reference deep-copies (src/Ciphertext.cpp:360-363).  Device buffers are never
    // modified in place once shared, so a copy shares them; += clones first when needed.
    if (o.certFHEcontext) certFHEcontext = new Context(*o.certFHEcontext);
    sharded = o.sharded;
    if (o.staged) {
        host_v = copy_words(o.host_v, o.host_len);
        if (o.host_bitlen) host_bitlen = copy_words(o.host_bitlen, o.host_len);
        host_len = o.host_len;
        host_v_valid = true;
        staged = true;
    } else {
        dev = o.dev;
        factors = o.factors;
        product_slot = o.product_slot;
    }
}

Ciphertext::Ciphertext(Ciphertext &&o) noexcept
    : dev(std::move(o.dev)), factors(std::move(o.factors)), product_slot(std::move(o.product_slot)),
      certFHEcontext(o.certFHEcontext), host_v(o.host_v),
      host_bitlen(o.host_bitlen), host_len(o.host_len), host_v_valid(o.host_v_valid), staged(o.staged),
      sharded(o.sharded) {
    o.factors.clear();
    o.product_slot.reset();
    o.certFHEcontext = nullptr;
    o.host_v = o.host_bitlen = nullptr;
    o.host_len = 0;
    o.host_v_valid = o.staged = false;
}

Ciphertext::~Ciphertext() {
    release();
    delete certFHEcontext;
    certFHEcontext = nullptr;
}

void Ciphertext::invalidate_mirror() const {
    delete[] host_v;
    delete[] host_bitlen;
    host_v = host_bitlen = nullptr;
    host_len = 0;
    host_v_valid = false;
}

void Ciphertext::release() {
    dev.reset();
    factors.clear();
    product_slot.reset();
    invalidate_mirror();
    staged = false;
}

void Ciphertext::upload_staged() {
    if (!staged) return;
    if (!certFHEcontext) throw Error("Ciphertext: values were set but no Context (call setContext first)");
    const uint64_t L = certFHEcontext->getDefaultN();
    if (L == 0 || host_len % L != 0)
        throw Error("Ciphertext: len = " + to_string(host_len) + " is not a multiple of the " + to_string(L) +
                    " words per block");
    if (host_bitlen) require_canonical_bitlen(host_bitlen, host_len, *certFHEcontext);
    glue::ensure_engine();
    csgn_buf *b = nullptr;
    glue::check(csgn_buf_upload_copy(host_v, host_len / L, (uint32_t)L, &b), "csgn_buf_upload_copy");
    dev = adopt(b);
    staged = false;  // host_v stays behind as a valid mirror
    host_v_valid = true;
}

bool Ciphertext::adopt_written_product() const {
    if (!product_slot || !*product_slot) return false;
    dev = *product_slot;
    factors.clear();
    product_slot.reset();
    return true;
}

void Ciphertext::materialize() const {
    // multiply the pending factors out, left to right: ((f0*f1)*f2)... -- the reference's own
    // i-major block order, whatever the grouping (block order of a product is associative)
    if (factors.empty() || adopt_written_product()) return;
    std::shared_ptr<csgn_buf> acc = factors[0];
    for (size_t i = 1; i < factors.size(); ++i) {
        csgn_buf *prod = nullptr;
        glue::check(csgn_mul(acc.get(), factors[i].get(), &prod), "csgn_mul");
        acc = adopt(prod);
    }
    dev = acc;
    factors.clear();
    if (product_slot) *product_slot = dev;
    product_slot.reset();
}

csgn_result *Ciphertext::materialize_and_fold(const csgn_key *key, int *status) const {
    // ((f0*f1)*...)*f_last with the LAST multiply fused with the decrypt fold: the kernel that writes the product
    // also counts its satisfied blocks (reference pattern tests/basic_operations.cpp:35-40; loops
    // src/Ciphertext.cpp:153-163 + src/SecretKey.cpp:131-140)
    std::shared_ptr<csgn_buf> acc = factors[0];
    for (size_t i = 1; i + 1 < factors.size(); ++i) {
        csgn_buf *prod = nullptr;
        glue::check(csgn_mul(acc.get(), factors[i].get(), &prod), "csgn_mul");
        acc = adopt(prod);
    }
    csgn_buf *prod = nullptr;
    csgn_result *res = nullptr;
    *status = csgn_mul_decrypt_deferred(acc.get(), factors.back().get(), key, &prod, &res);
    if (*status != CSGN_OK) return nullptr;
    dev = adopt(prod);
    factors.clear();
    if (product_slot) *product_slot = dev;
    product_slot.reset();
    return res;
}

void Ciphertext::collect_factors(std::vector<std::shared_ptr<csgn_buf> > &into, bool flatten) const {
    // flatten (lazy mode): the factors of a pending product become factors of the new one.  Otherwise (fused mode)
    // an operand that is itself a pending product is written now -- once, whoever uses it afterwards -- so that a
    // pending product always has exactly two written factors.
    const_cast<Ciphertext *>(this)->upload_staged();
    if (!flatten) materialize();
    if (!factors.empty()) into.insert(into.end(), factors.begin(), factors.end());
    else if (dev) into.push_back(dev);
}

void Ciphertext::setValues(const uint64_t *V, const uint64_t length) {
    // reference src/Ciphertext.cpp:392-403: replaces v and len; bitlen is left alone
    uint64_t *fresh = copy_words(V, length);
    uint64_t *keep_bitlen = (host_bitlen && host_len == length && staged) ? host_bitlen : nullptr;
    if (keep_bitlen) host_bitlen = nullptr;
    release();
    host_v = fresh;
    host_bitlen = keep_bitlen;
    host_len = length;
    host_v_valid = true;
    staged = true;
    if (certFHEcontext && certFHEcontext->getDefaultN() && length % certFHEcontext->getDefaultN() == 0) upload_staged();
}

void Ciphertext::setBitlen(const uint64_t *Bitlen, const uint64_t length) {
    // reference src/Ciphertext.cpp:405-415.  Only the canonical pattern is representable.
    if (certFHEcontext) {
        require_canonical_bitlen(Bitlen, length, *certFHEcontext);
        if (length != getLen())
            throw Error("Ciphertext::setBitlen: length " + to_string(length) + " differs from the " +
                        to_string(getLen()) + " words held");
        return;
    }
    // no context yet: remember it, it is checked when the words go to the device
    uint64_t *fresh = copy_words(Bitlen, length);
    delete[] host_bitlen;
    host_bitlen = fresh;
    if (!staged) {
        staged = true;
        host_len = length;
        delete[] host_v;
        host_v = new uint64_t[length ? length : 1]();
        host_v_valid = true;
    }
}

void Ciphertext::setContext(const Context &context) {
    Context *fresh = new Context(context);
    delete certFHEcontext;
    certFHEcontext = fresh;
    if (staged && fresh->getDefaultN() && host_len % fresh->getDefaultN() == 0) upload_staged();
}

Ciphertext Ciphertext::shard() const {
    if (sharded) return *this;
    const csgn_buf *whole = deviceBuffer();
    if (!whole) throw Error("Ciphertext::shard: empty ciphertext");
    uint64_t first = 0, count = 0;
    glue::check(csgn_shard_range(csgn_buf_blocks(whole), Library::getRank(), Library::getWorldSize(), &first, &count),
                "csgn_shard_range");
    csgn_buf *mine = nullptr;
    glue::check(csgn_buf_slice(whole, first, count, &mine), "csgn_buf_slice");
    Ciphertext out;
    if (certFHEcontext) out.certFHEcontext = new Context(*certFHEcontext);
    out.dev = adopt(mine);
    out.sharded = true;
    return out;
}

bool Ciphertext::isSharded() const { return sharded; }

uint64_t Ciphertext::getBlocks() const {
    if (staged) {
        const uint64_t L = certFHEcontext ? certFHEcontext->getDefaultN() : 0;
        return L ? host_len / L : 0;
    }
    if (!factors.empty()) {   // pending product: the block counts multiply (saturating)
        uint64_t n = 1;
        for (size_t i = 0; i < factors.size(); ++i) {
            const uint64_t t = csgn_buf_blocks(factors[i].get());
            n = (t != 0 && n > UINT64_MAX / t) ? UINT64_MAX : n * t;
        }
        return n;
    }
    return dev ? csgn_buf_blocks(dev.get()) : 0;
}

uint64_t Ciphertext::getLen() const {
    if (staged) return host_len;
    const csgn_buf *any = !factors.empty() ? factors[0].get() : dev.get();
    if (!any) return 0;
    const uint64_t n = getBlocks(), L = csgn_buf_words_per_block(any);
    return (n > UINT64_MAX / L) ? UINT64_MAX : n * L;
}

Context Ciphertext::getContext() const {
    if (!certFHEcontext) throw Error("Ciphertext::getContext: this ciphertext has no Context");
    return *certFHEcontext;
}

const csgn_buf *Ciphertext::deviceBuffer() const {
    const_cast<Ciphertext *>(this)->upload_staged();
    materialize();
    return dev.get();
}

uint64_t *Ciphertext::getValues() const {
    if (staged) return host_v;
    const csgn_buf *buf = deviceBuffer();
    if (!buf) return nullptr;
    const uint64_t len = getLen();
    if (!host_v_valid || host_len != len) {
        delete[] host_v;
        host_v = new uint64_t[len ? len : 1];
        if (host_len != len) {
            delete[] host_bitlen;
            host_bitlen = nullptr;
        }
        host_len = len;
        glue::check(csgn_buf_download(buf, host_v), "csgn_buf_download");
        host_v_valid = true;
    }
    return host_v;
}

uint64_t *Ciphertext::getBitlen() const {
    if (staged && !certFHEcontext) return host_bitlen;
    if (!factors.empty()) materialize();   // the caller is about to index len words
    const uint64_t len = getLen();
    if (!len || !certFHEcontext) return nullptr;
    if (!host_bitlen || host_len != len) {
        if (host_len != len && !staged) {  // the word mirror is for another length
            delete[] host_v;
            host_v = nullptr;
            host_v_valid = false;
        }
        delete[] host_bitlen;
        host_bitlen = new uint64_t[len];
        host_len = len;
        const uint64_t L = certFHEcontext->getDefaultN(), rem = certFHEcontext->getN() % 64;
        for (uint64_t i = 0; i < len; ++i) host_bitlen[i] = canonical_bits(i, L, rem);
    }
    return host_bitlen;
}

ostream &operator<<(ostream &out, const Ciphertext &c) {
    // reference src/Ciphertext.cpp:185-202: the valid bits of every word, MSB first
    const uint64_t *v = c.getValues();
    const uint64_t *bl = c.getBitlen();
    const uint64_t len = c.getLen();
    for (uint64_t i = 0; i < len; ++i)
        for (uint64_t k = 0; k < (bl ? bl[i] : 64); ++k) out << ((v[i] >> (63 - k)) & 1ull);
    out << endl;
    return out;
}

long Ciphertext::size() {
    // reference src/Ciphertext.cpp:91-101: three pointers, one length, two arrays of len words
    return (long)(sizeof(void *) * 3 + sizeof(uint64_t) + getLen() * 2 * sizeof(uint64_t));
}

Ciphertext &Ciphertext::operator=(const Ciphertext &c) {
    // the reference frees the context here and never restores it (src/Ciphertext.cpp:306-329);
    // this is a complete copy
    if (this == &c) return *this;
    Ciphertext tmp(c);
    *this = std::move(tmp);
    return *this;
}

Ciphertext &Ciphertext::operator=(Ciphertext &&o) noexcept {
    if (this == &o) return *this;
    release();
    delete certFHEcontext;
    dev = std::move(o.dev);
    factors = std::move(o.factors);
    o.factors.clear();
    product_slot = std::move(o.product_slot);
    o.product_slot.reset();
    certFHEcontext = o.certFHEcontext;
    host_v = o.host_v;
    host_bitlen = o.host_bitlen;
    host_len = o.host_len;
    host_v_valid = o.host_v_valid;
    staged = o.staged;
    sharded = o.sharded;
    o.certFHEcontext = nullptr;
    o.host_v = o.host_bitlen = nullptr;
    o.host_len = 0;
    o.host_v_valid = o.staged = false;
    return *this;
}

namespace {
// a + b over several GPUs: both sharded (each rank concatenates its parts; block ORDER across ranks is then no
// longer the single-process order, which decrypt does not observe) or neither
void require_same_sharding(bool a, bool b, bool a_empty, bool b_empty, const char *op) {
    if (a != b && !a_empty && !b_empty)
        throw Error(string(op) + ": one operand is sharded and the other replicated (the replicated blocks would be "
                    "counted once per rank); call shard() on it first");
}
}  // namespace

Ciphertext Ciphertext::operator+(const Ciphertext &c) const {
    const csgn_buf *a = deviceBuffer(), *b = c.deviceBuffer();
    require_same_sharding(sharded, c.sharded, !a, !b, "Ciphertext::operator+");
    Ciphertext out;
    out.sharded = (a && sharded) || (b && c.sharded);
    const Context *ctx = certFHEcontext ? certFHEcontext : c.certFHEcontext;
    if (ctx) out.certFHEcontext = new Context(*ctx);
    if (a && b) {
        csgn_buf *sum = nullptr;
        // the reference copies both operands (src/Ciphertext.cpp:215-223, then again in the constructor); large
        // operands are referred to instead (a lazy sum), small ones copied once
        if (Library::getRopeSums()) glue::check(csgn_concat_lazy(a, b, &sum), "csgn_concat_lazy");
        else glue::check(csgn_concat(a, b, &sum), "csgn_concat");
        out.dev = adopt(sum);
    } else if (a || b) {
        out.dev = a ? dev : c.dev;   // x + (empty) is x: share it
    }
    return out;
}

Ciphertext &Ciphertext::operator+=(const Ciphertext &c) {
    std::shared_ptr<csgn_buf> rhs;   // keep the operand alive (and unchanged) when &c == this
    {
        c.deviceBuffer();
        rhs = c.dev;
    }
    deviceBuffer();
    if (!rhs) return *this;
    require_same_sharding(sharded, c.sharded, !dev, false, "Ciphertext::operator+=");
    if (!certFHEcontext && c.certFHEcontext) certFHEcontext = new Context(*c.certFHEcontext);
    if (!dev) {
        sharded = c.sharded;
        dev = rhs;
    } else {
        if (dev.use_count() > 1 + (dev == rhs ? 1 : 0) || csgn_buf_retained(dev.get())) {
            // shared with a copy, or part of a lazy sum: do not grow it under the other owner
            csgn_buf *mine = nullptr;
            glue::check(csgn_buf_clone(dev.get(), &mine), "csgn_buf_clone");
            dev = adopt(mine);
        }
        glue::check(csgn_append(dev.get(), rhs.get()), "csgn_append");
    }
    invalidate_mirror();
    return *this;
}

Ciphertext Ciphertext::operator*(const Ciphertext &c) const {
    if (c.sharded)
        throw Error("Ciphertext::operator*: the right operand of a product must be replicated (the LEFT operand is the "
                    "sharded one: rank g then owns output blocks [first*T2, (first+count)*T2), SURVEY.md 8e)");
    Ciphertext out;
    out.sharded = sharded;
    // the reference multiplies in the LEFT operand's context (src/Ciphertext.cpp:239)
    if (certFHEcontext) out.certFHEcontext = new Context(*certFHEcontext);
    if (Library::getLazyProducts() || Library::getFusedProducts()) {
        // the product is not written yet: decrypt folds it in the kernel that writes it (fused) or never writes it (lazy)
        const bool flatten = Library::getLazyProducts();
        collect_factors(out.factors, flatten);
        const size_t mine = out.factors.size();
        c.collect_factors(out.factors, flatten);
        if (mine == 0 || out.factors.size() == mine) throw Error("Ciphertext::operator*: empty operand");
        out.product_slot = std::make_shared<std::shared_ptr<csgn_buf> >();
        return out;
    }
    const csgn_buf *a = deviceBuffer(), *b = c.deviceBuffer();
    if (!a || !b) throw Error("Ciphertext::operator*: empty operand");
    csgn_buf *prod = nullptr;
    glue::check(csgn_mul(a, b, &prod), "csgn_mul");
    out.dev = adopt(prod);
    return out;
}

Ciphertext &Ciphertext::operator*=(const Ciphertext &c) {
    if (c.sharded) throw Error("Ciphertext::operator*=: the right operand of a product must be replicated");
    if (Library::getLazyProducts() || Library::getFusedProducts()) {
        const bool flatten = Library::getLazyProducts();
        std::vector<std::shared_ptr<csgn_buf> > f;
        collect_factors(f, flatten);
        const size_t mine = f.size();
        c.collect_factors(f, flatten);
        if (mine == 0 || f.size() == mine) throw Error("Ciphertext::operator*=: empty operand");
        dev.reset();
        factors.swap(f);
        product_slot = std::make_shared<std::shared_ptr<csgn_buf> >();
        invalidate_mirror();
        return *this;
    }
    const csgn_buf *a = deviceBuffer(), *b = c.deviceBuffer();
    if (!a || !b) throw Error("Ciphertext::operator*=: empty operand");
    csgn_buf *prod = nullptr;
    glue::check(csgn_mul(a, b, &prod), "csgn_mul");
    dev = adopt(prod);   // the old buffer is released in stream order, after the multiply that reads it
    invalidate_mirror();
    return *this;
}

void Ciphertext::save(const std::string &path) const {
    const csgn_buf *buf = deviceBuffer();
    if (!buf || !certFHEcontext) throw Error("Ciphertext::save: empty ciphertext or no Context");
    glue::check(csgn_buf_save(buf, certFHEcontext->getN(), certFHEcontext->getD(), path.c_str()), "csgn_buf_save");
}

Ciphertext Ciphertext::load(const std::string &path) {
    glue::ensure_engine();
    uint64_t n = 0, d = 0;
    csgn_buf *buf = nullptr;
    glue::check(csgn_buf_load(path.c_str(), &n, &d, &buf), "csgn_buf_load");
    Ciphertext out;
    out.dev = adopt(buf);
    out.certFHEcontext = new Context(n, d);
    return out;
}

void Ciphertext::saveSharded(const std::string &prefix) const {
    const csgn_buf *buf = deviceBuffer();
    if (!buf || !certFHEcontext) throw Error("Ciphertext::saveSharded: empty ciphertext or no Context");
    glue::check(csgn_buf_save_shard(buf, certFHEcontext->getN(), certFHEcontext->getD(), prefix.c_str(), Library::getRank(),
                                    Library::getWorldSize(), 0),
                "csgn_buf_save_shard");
}

Ciphertext Ciphertext::loadSharded(const std::string &prefix) {
    glue::ensure_engine();
    uint64_t n = 0, d = 0;
    csgn_buf *buf = nullptr;
    glue::check(csgn_buf_load_shard(prefix.c_str(), Library::getRank(), Library::getWorldSize(), &n, &d, nullptr, &buf),
                "csgn_buf_load_shard");
    Ciphertext out;
    out.dev = adopt(buf);
    out.certFHEcontext = new Context(n, d);
    out.sharded = true;
    return out;
}

void Ciphertext::applyPermutation_inplace(const Permutation &permutation) {
    Ciphertext permuted = applyPermutation(permutation);
    dev = std::move(permuted.dev);
    factors.swap(permuted.factors);
    product_slot = std::move(permuted.product_slot);
    invalidate_mirror();
}

Ciphertext Ciphertext::applyPermutation(const Permutation &permutation) {
    upload_staged();
    if (!certFHEcontext || (!dev && factors.empty()))
        throw Error("Ciphertext::applyPermutation: empty ciphertext or no Context");
    const csgn_perm *map = permutation.deviceMap(certFHEcontext->getN());
    const bool strict = Library::getStrictReferencePermutation();
    Ciphertext out;
    out.certFHEcontext = new Context(*certFHEcontext);
    out.sharded = sharded;
    if (!factors.empty()) adopt_written_product();
    if (!factors.empty() && !strict) {
        // a permutation acts on each block, and a product block is an AND of factor blocks:
        // pi(a_i & b_j) = pi(a_i) & pi(b_j) -- permute the factors, the product stays pending
        out.product_slot = std::make_shared<std::shared_ptr<csgn_buf> >();
        for (size_t i = 0; i < factors.size(); ++i) {
            csgn_buf *pf = nullptr;
            glue::check(csgn_permute(factors[i].get(), map, 0, &pf), "csgn_permute");
            out.factors.push_back(adopt(pf));
        }
        return out;
    }
    csgn_buf *res = nullptr;
    glue::check(csgn_permute(deviceBuffer(), map, strict ? 1 : 0, &res), "csgn_permute");
    out.dev = adopt(res);
    return out;
}

// =========================================================================================
// SecretKey
// =========================================================================================
SecretKey::SecretKey(const Context &context) : s(nullptr), length(0), certFHEContext(nullptr), device_key(nullptr) {
    // reference src/SecretKey.cpp:308-337: reseed from the clock, then rejection-sample D
    // distinct positions.  The reference compares against slots it has not written yet
    // (uninitialised reads); the slots start at an impossible value here instead.
    srand(time(NULL));
    certFHEContext = new Context(context);
    const uint64_t d = context.getD(), n = context.getN();
    s = new uint64_t[d ? d : 1];
    length = (long)d;
    for (uint64_t i = 0; i < d; ++i) s[i] = (uint64_t)-1;
    if (d > n) throw Error("SecretKey: D = " + to_string(d) + " secret positions do not fit in N = " + to_string(n));
    uint64_t count = 0;
    while (count < d) {
        const uint64_t t = (uint64_t)rand() % n;
        if (Helper::exists(s, d, t)) continue;
        s[count++] = t;
    }
}

SecretKey::SecretKey(const SecretKey &k) : s(nullptr), length(0), certFHEContext(nullptr), device_key(nullptr) {
    certFHEContext = new Context(*k.certFHEContext);
    length = k.length < 0 ? 0 : k.length;
    s = copy_words(k.s, (uint64_t)length);
}

SecretKey::~SecretKey() {
    drop_device_key();
    // the reference zeroises the positions before freeing them (src/SecretKey.cpp:356-357)
    for (long i = 0; i < length; ++i) ((volatile uint64_t *)s)[i] = 0;
    delete[] s;
    s = nullptr;
    length = -1;
    delete certFHEContext;
    certFHEContext = nullptr;
}

void SecretKey::drop_device_key() const {
    if (device_key) {
        csgn_key_free(device_key);  // zeroises the device mask
        device_key = nullptr;
    }
}

SecretKey &SecretKey::operator=(const SecretKey &k) {
    if (this == &k) return *this;
    drop_device_key();
    uint64_t *fresh = copy_words(k.s, (uint64_t)(k.length < 0 ? 0 : k.length));
    for (long i = 0; i < length; ++i) ((volatile uint64_t *)s)[i] = 0;
    delete[] s;
    s = fresh;
    length = k.length < 0 ? 0 : k.length;
    *certFHEContext = *k.certFHEContext;
    return *this;
}

ostream &operator<<(ostream &out, const SecretKey &c) {
    // reference src/SecretKey.cpp:22-29
    for (uint64_t i = 0; i < c.getLength(); ++i) out << c.getKey()[i] << " ";
    out << endl;
    return out;
}

uint64_t SecretKey::getLength() const { return (uint64_t)length; }
uint64_t *SecretKey::getKey() const { return s; }

void SecretKey::setKey(uint64_t *key, uint64_t len) {
    drop_device_key();
    uint64_t *fresh = copy_words(key, len);
    delete[] s;
    s = fresh;
    length = (long)len;
}

long SecretKey::size() {
    // reference src/SecretKey.cpp:269-276
    return (long)(sizeof(void *) + sizeof(long) + sizeof(uint64_t) * (uint64_t)length);
}

void SecretKey::save(const std::string &path) const {
    glue::check(csgn_key_positions_save(path.c_str(), certFHEContext->getN(), certFHEContext->getD(), s, (uint64_t)length),
                "csgn_key_positions_save");
}

SecretKey SecretKey::load(const std::string &path) {
    uint64_t n = 0, d = 0, count = 0;
    glue::check(csgn_key_positions_load(path.c_str(), &n, &d, nullptr, 0, &count), "csgn_key_positions_load");
    std::vector<uint64_t> pos(count ? count : 1);
    glue::check(csgn_key_positions_load(path.c_str(), &n, &d, pos.data(), count, &count), "csgn_key_positions_load");
    // the constructor reseeds rand() from the clock and draws a key (src/SecretKey.cpp:308-337): let it do that on a
    // scratch generator state, so that the caller's rand() sequence continues exactly where it was
    char scratch[128];
    char *caller_state = initstate(1u, scratch, sizeof scratch);
    SecretKey key((Context(n, d)));
    setstate(caller_state);
    key.setKey(pos.data(), count);
    for (size_t i = 0; i < pos.size(); ++i) ((volatile uint64_t *)pos.data())[i] = 0;
    return key;
}

uint64_t *SecretKey::encrypt(unsigned char bit, uint64_t n, uint64_t d, uint64_t *key) {
    // One value (0/1) per position, drawn in the reference's order (src/SecretKey.cpp:35-80).
    // The reference asks Helper::exists(key, d, i) for every position (O(N*D)); an indicator
    // vector answers the same question, with the same rand() calls in the same order.
    uint64_t *res = new uint64_t[n ? n : 1];
    std::vector<unsigned char> secret(n ? n : 1, 0);
    for (uint64_t k = 0; k < d; ++k)
        if (key[k] < n) secret[key[k]] = 1;
    if (bit & 0x01) {
        // Enc(1): secret positions are 1, one rand() for every other position, in index order
        for (uint64_t i = 0; i < n; ++i) res[i] = secret[i] ? 1 : (uint64_t)(rand() % 2);
        return res;
    }
    // Enc(0): pick the secret position that may break the all-ones pattern, randomise the
    // rest, and force a zero there only if every other secret position came out 1
    const uint64_t hole = key[(uint64_t)rand() % d];
    uint64_t others = 0;
    bool seen = false;
    for (uint64_t i = 0; i < n; ++i) {
        if (i == hole) continue;
        res[i] = (uint64_t)(rand() % 2);
        if (secret[i]) {
            others = seen ? (others & res[i]) : res[i];
            seen = true;
        }
    }
    res[hole] = (others == 1) ? 0 : (uint64_t)(rand() % 2);
    return res;
}

Ciphertext SecretKey::encrypt(Plaintext &plaintext) {
    // reference src/SecretKey.cpp:153-206: per-bit values, then MSB-first packing
    const uint64_t n = certFHEContext->getN(), d = certFHEContext->getD();
    const uint64_t L = certFHEContext->getDefaultN();
    if ((uint64_t)length < d) throw Error("SecretKey::encrypt: the key holds fewer positions than D");
    uint64_t *bits = encrypt(plaintext.getValue(), n, d, s);
    uint64_t *words = new uint64_t[L ? L : 1]();
    for (uint64_t p = 0; p < n; ++p) words[p >> 6] |= (bits[p] & 1ull) << (63 - (p & 63));
    delete[] bits;
    Ciphertext c(words, nullptr, L, *certFHEContext);
    delete[] words;
    return c;
}

Ciphertext SecretKey::encryptBatch(const unsigned char *bits, uint64_t n, uint64_t seed) {
    if (!device_key) {
        glue::ensure_engine();
        glue::check(csgn_key_create(certFHEContext->getN(), s, (uint32_t)length, &device_key), "csgn_key_create");
    }
    csgn_buf *buf = nullptr;
    glue::check(csgn_encrypt_batch(device_key, bits, n, 0, seed, &buf), "csgn_encrypt_batch");
    Ciphertext out;
    out.dev = adopt(buf);
    out.certFHEcontext = new Context(*certFHEContext);
    return out;
}

void SecretKey::decryptBatch(Ciphertext *ciphertexts, uint64_t n, unsigned char *bits) {
    if (n == 0) return;
    if (!ciphertexts || !bits) throw Error("SecretKey::decryptBatch: null argument");
    if (!device_key) {
        glue::ensure_engine();
        glue::check(csgn_key_create(certFHEContext->getN(), s, (uint32_t)length, &device_key), "csgn_key_create");
    }
    std::vector<const csgn_buf *> bufs;
    std::vector<uint64_t> where;
    std::vector<std::pair<uint64_t, Plaintext> > later;
    for (uint64_t i = 0; i < n; ++i) {
        Ciphertext &c = ciphertexts[i];
        c.upload_staged();
        if (c.sharded || !c.factors.empty() || (!c.dev && c.factors.empty())) {
            // sharded, pending or empty: the single-ciphertext path knows how; its value is read after everything
            // has been enqueued
            later.push_back(std::make_pair(i, decrypt(c)));
            continue;
        }
        bufs.push_back(c.dev.get());
        where.push_back(i);
    }
    if (!bufs.empty()) {
        std::vector<uint8_t> out(bufs.size());
        glue::check(csgn_decrypt_batch(bufs.data(), (uint32_t)bufs.size(), device_key, out.data(), nullptr), "csgn_decrypt_batch");
        for (size_t k = 0; k < bufs.size(); ++k) bits[where[k]] = out[k];
    }
    for (size_t k = 0; k < later.size(); ++k) bits[later[k].first] = later[k].second.getValue();
}

namespace {
// a Plaintext whose value is read from `res` on first use
std::shared_ptr<csgn_result> adopt_result(csgn_result *r) {
    return std::shared_ptr<csgn_result>(r, [](csgn_result *p) { csgn_result_free(p); });
}
}  // namespace

Plaintext SecretKey::decrypt(Ciphertext &ciphertext) {
    ciphertext.upload_staged();
    if (!ciphertext.dev && ciphertext.factors.empty()) return Plaintext(0);
    if (!device_key) {
        glue::ensure_engine();
        glue::check(csgn_key_create(certFHEContext->getN(), s, (uint32_t)length, &device_key), "csgn_key_create");
    }
    uint8_t bit = 0;
    if (ciphertext.sharded) {
        // this rank's blocks only: the fold kernel publishes its count to the peers and collects theirs
        if (!glue::comm()) throw Error("SecretKey::decrypt: sharded ciphertext but no peers (Library::connectPeers)");
        const csgn_buf *local = ciphertext.deviceBuffer();   // multiplies pending factors out
        if (!local) throw Error("SecretKey::decrypt: sharded ciphertext without blocks");
        glue::check(csgn_decrypt_sharded(local, device_key, glue::comm(), &bit, nullptr), "csgn_decrypt_sharded");
        return Plaintext(bit);
    }
    if (!ciphertext.factors.empty() && Library::getLazyProducts()) {
        // a product that is never multiplied out: Dec(f0*f1*...) = Dec(f0) & Dec(f1) & ...
        std::vector<const csgn_buf *> fs;
        for (size_t i = 0; i < ciphertext.factors.size(); ++i) fs.push_back(ciphertext.factors[i].get());
        glue::check(csgn_decrypt_product(fs.data(), (uint32_t)fs.size(), device_key, &bit, nullptr), "csgn_decrypt_product");
        return Plaintext(bit);
    }
    // The fold (fused with the multiply when the product has not been written yet) and the copy of its count are
    // enqueued; the Plaintext reads the count on first use.  When every result slot is taken (thousands of unread
    // Plaintexts) the call degrades to the blocking form.
    csgn_result *res = nullptr;
    if (!ciphertext.factors.empty()) ciphertext.adopt_written_product();
    if (ciphertext.factors.size() >= 2) {
        int rc = CSGN_OK;
        res = ciphertext.materialize_and_fold(device_key, &rc);
        if (rc != CSGN_OK && rc != CSGN_ERR_OUT_OF_MEMORY) glue::check(rc, "csgn_mul_decrypt_deferred");
    }
    if (!res) {
        const csgn_buf *buf = ciphertext.deviceBuffer();
        const int rc = csgn_decrypt_deferred(buf, device_key, &res);
        if (rc == CSGN_ERR_OUT_OF_MEMORY) {
            glue::check(csgn_decrypt(buf, device_key, &bit), "csgn_decrypt");
            return Plaintext(bit);
        }
        glue::check(rc, "csgn_decrypt_deferred");
    }
    Plaintext out;
    out.pending = adopt_result(res);
    return out;
}

void SecretKey::applyPermutation_inplace(const Permutation &permutation) {
    // reference src/SecretKey.cpp:226-259: new key = ascending { i : perm[i] is a secret position }
    const uint64_t n = certFHEContext->getN();
    if (permutation.getLength() != n)
        throw Error("SecretKey::applyPermutation: permutation length differs from N");
    const uint64_t *perm = permutation.getPermutation();
    vector<unsigned char> secret(n, 0);
    for (long i = 0; i < length; ++i) secret[s[i]] = 1;
    uint64_t *fresh = new uint64_t[length > 0 ? length : 1];
    long out = 0;
    for (uint64_t i = 0; i < n && out < length; ++i)
        if (secret[perm[i]]) fresh[out++] = i;
    drop_device_key();
    for (long i = 0; i < length; ++i) ((volatile uint64_t *)s)[i] = 0;
    delete[] s;
    s = fresh;
}

SecretKey SecretKey::applyPermutation(const Permutation &permutation) {
    SecretKey k(*this);
    k.applyPermutation_inplace(permutation);
    return k;
}

}  // namespace certFHE
