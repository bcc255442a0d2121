// diff_vs_reference.cpp -- one process, two libraries: the certFHE drop-in of this
// repository (namespace certFHE, GPU engine) and the UNMODIFIED reference build
// (oracle/_ref/libcertfhe_ref.so, compiled with -DcertFHE=certFHE_ref).  Both are
// driven through the same public class API on identical seeded inputs and every
// observable -- words, bitlen, lengths, size(), decrypted bits, printed text,
// permutations, permuted keys -- must agree bit for bit.
//
// Built only where /root/reference is present (the reference HEADERS are needed to
// compile this file; nothing of the reference is copied into the repository).  The
// binary travels to the GPU box.  The reference's undefined behaviour is avoided:
// no operator= followed by multiply, no N % 64 == 0, no empty decrypt.
#include "certFHE.h"  // ours

#define certFHE certFHE_ref
#include CSGN_REFERENCE_HEADER  // "/root/reference/src/certFHE.h"
#undef certFHE

#include <chrono>
#include <random>
#include <sstream>

namespace ours = certFHE;
namespace ref = certFHE_ref;

static int failures = 0;
#define EXPECT(cond)                                                                          \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            std::cerr << "FAILED " << __FILE__ << ":" << __LINE__ << "  " #cond << std::endl; \
            ++failures;                                                                       \
        }                                                                                     \
    } while (0)

template <class A, class B>
static bool same_ct(A &a, B &b) {
    if (a.getLen() != b.getLen()) return false;
    if (a.size() != b.size()) return false;
    const uint64_t n = a.getLen();
    if (n == 0) return true;
    return memcmp(a.getValues(), b.getValues(), n * 8) == 0 && memcmp(a.getBitlen(), b.getBitlen(), n * 8) == 0;
}

template <class T>
static std::string text(const T &x) {
    std::ostringstream os;
    os << x;
    return os.str();
}

static std::vector<uint64_t> seeded_key(uint64_t N, uint64_t D, uint64_t seed) {
    std::mt19937_64 g(seed);
    std::vector<uint64_t> pool(N);
    for (uint64_t i = 0; i < N; ++i) pool[i] = i;
    for (uint64_t i = 0; i < D; ++i) std::swap(pool[i], pool[i + g() % (N - i)]);
    pool.resize(D);
    return pool;
}

static void encrypted_circuits(uint64_t N, uint64_t D, unsigned seed, int n_a, int n_b) {
    ours::Context octx(N, D);
    ref::Context rctx(N, D);
    EXPECT(text(octx) == text(rctx));
    EXPECT(octx.getS() == rctx.getS() && octx.getDefaultN() == rctx.getDefaultN());
    ours::SecretKey osk(octx);
    ref::SecretKey rsk(rctx);
    std::vector<uint64_t> key = seeded_key(N, D, seed);
    osk.setKey(key.data(), D);
    rsk.setKey(key.data(), D);
    EXPECT(text(osk) == text(rsk) && osk.size() == rsk.size() && osk.getLength() == rsk.getLength());

    // identical rand() streams -> identical fresh ciphertexts
    std::vector<int> bits_a(n_a), bits_b(n_b);
    std::mt19937 pick(seed);
    for (int &b : bits_a) b = pick() & 1;
    for (int &b : bits_b) b = pick() & 1;
    srand(seed);
    std::vector<ref::Ciphertext> ra, rb;
    for (int b : bits_a) { ref::Plaintext p(b); ra.push_back(rsk.encrypt(p)); }
    for (int b : bits_b) { ref::Plaintext p(b); rb.push_back(rsk.encrypt(p)); }
    srand(seed);
    std::vector<ours::Ciphertext> oa, ob;
    for (int b : bits_a) { ours::Plaintext p(b); oa.push_back(osk.encrypt(p)); }
    for (int b : bits_b) { ours::Plaintext p(b); ob.push_back(osk.encrypt(p)); }
    for (int i = 0; i < n_a; ++i) EXPECT(same_ct(oa[i], ra[i]));
    for (int j = 0; j < n_b; ++j) EXPECT(same_ct(ob[j], rb[j]));
    if (N <= 256) EXPECT(text(oa[0]) == text(ra[0]));

    // sums via += and +
    ours::Ciphertext oA(oa[0]), oB(ob[0]);
    ref::Ciphertext rA(ra[0]), rB(rb[0]);
    for (int i = 1; i < n_a; ++i) { oA += oa[i]; rA += ra[i]; }
    for (int j = 1; j < n_b; ++j) {
        ours::Ciphertext t = oB + ob[j];
        ref::Ciphertext u = rB + rb[j];
        EXPECT(same_ct(t, u));
        oB += ob[j];
        rB += rb[j];
    }
    EXPECT(same_ct(oA, rA) && same_ct(oB, rB));

    // product via * and *=, then one more level
    ours::Ciphertext oP = oA * oB;
    ref::Ciphertext rP = rA * rB;
    EXPECT(same_ct(oP, rP));
    ours::Ciphertext oQ(oA);
    ref::Ciphertext rQ(rA);
    oQ *= oB;
    rQ *= rB;
    EXPECT(same_ct(oQ, rQ));
    ours::Ciphertext oR = oP * oa[0];
    ref::Ciphertext rR = rP * ra[0];
    EXPECT(same_ct(oR, rR));
    ours::Ciphertext oS = oR + oP;
    ref::Ciphertext rS = rR + rP;
    EXPECT(same_ct(oS, rS));

    // decrypt everything
    EXPECT(osk.decrypt(oA).getValue() == rsk.decrypt(rA).getValue());
    EXPECT(osk.decrypt(oB).getValue() == rsk.decrypt(rB).getValue());
    EXPECT(osk.decrypt(oP).getValue() == rsk.decrypt(rP).getValue());
    EXPECT(osk.decrypt(oR).getValue() == rsk.decrypt(rR).getValue());
    EXPECT(osk.decrypt(oS).getValue() == rsk.decrypt(rS).getValue());
    EXPECT(osk.decrypt(oa[0]).getValue() == bits_a[0] && rsk.decrypt(ra[0]).getValue() == bits_a[0]);
    ours::Plaintext op = osk.decrypt(oP);
    ref::Plaintext rp = rsk.decrypt(rP);
    EXPECT(text(op) == text(rp));

    // permutations: same rand() stream -> same permutation, inverse, composition, key
    srand(seed + 1);
    ref::Permutation rperm(rctx);
    srand(seed + 1);
    ours::Permutation operm(octx);
    EXPECT(operm.getLength() == rperm.getLength());
    EXPECT(memcmp(operm.getPermutation(), rperm.getPermutation(), N * 8) == 0);
    if (N <= 256) EXPECT(text(operm) == text(rperm));
    ours::Permutation oinv = operm.getInverse();
    ref::Permutation rinv = rperm.getInverse();
    EXPECT(memcmp(oinv.getPermutation(), rinv.getPermutation(), N * 8) == 0);
    srand(seed + 2);
    ref::Permutation rperm2(N);
    srand(seed + 2);
    ours::Permutation operm2(N);
    ours::Permutation ocomp = operm + operm2;
    ref::Permutation rcomp = rperm + rperm2;
    EXPECT(memcmp(ocomp.getPermutation(), rcomp.getPermutation(), N * 8) == 0);
    operm2 += operm;
    rperm2 += rperm;
    EXPECT(memcmp(operm2.getPermutation(), rperm2.getPermutation(), N * 8) == 0);
    ours::SecretKey opk = osk.applyPermutation(operm);
    ref::SecretKey rpk = rsk.applyPermutation(rperm);
    EXPECT(text(opk) == text(rpk));

    // single block: the reference permutes it correctly -> whole objects agree
    ours::Ciphertext op1 = oa[0].applyPermutation(operm);
    ref::Ciphertext rp1 = ra[0].applyPermutation(rperm);
    EXPECT(same_ct(op1, rp1));
    EXPECT(opk.decrypt(op1).getValue() == rpk.decrypt(rp1).getValue());
    // multi block, reference-strict mode: the reference keeps block 0 only (src/Ciphertext.cpp:33-40)
    ours::Library::setStrictReferencePermutation(true);
    ours::Ciphertext ostrict = oP.applyPermutation(operm);
    ours::Ciphertext ostrict2(oP);
    ostrict2.applyPermutation_inplace(operm);
    ours::Library::setStrictReferencePermutation(false);
    ref::Ciphertext rstrict = rP.applyPermutation(rperm);
    EXPECT(same_ct(ostrict, rstrict) && same_ct(ostrict2, rstrict));
    // multi block, all-blocks mode: equals the reference applied to each block on its own
    ours::Ciphertext oall = oP.applyPermutation(operm);
    EXPECT(oall.getLen() == oP.getLen());
    const uint64_t L = octx.getDefaultN();
    std::vector<uint64_t> bl(L, 64);
    if (N % 64) bl[L - 1] = N % 64;
    for (uint64_t k = 0; k < oP.getLen() / L; ++k) {
        ref::Ciphertext one(rP.getValues() + k * L, bl.data(), L, rctx);
        ref::Ciphertext onep = one.applyPermutation(rperm);
        EXPECT(memcmp(oall.getValues() + k * L, onep.getValues(), L * 8) == 0);
    }
    EXPECT(opk.decrypt(oall).getValue() == osk.decrypt(oP).getValue());
}

// BASELINE.json configs[1]: two 1,000-block ciphertexts -> 1,000,000 output blocks, then decrypt.
// Raw seeded blocks with the key bits planted in a few of them; every output word is compared.
static void config2_full_size(uint64_t N, uint64_t D, uint64_t T1, uint64_t T2) {
    ours::Context octx(N, D);
    ref::Context rctx(N, D);
    const uint64_t L = octx.getDefaultN(), rem = N % 64;
    const uint64_t pad = rem ? ~0ull << (64 - rem) : ~0ull;
    std::vector<uint64_t> key = seeded_key(N, D, 77);
    std::vector<uint64_t> mask(L, 0);
    for (uint64_t s : key) mask[s >> 6] |= 1ull << (63 - (s & 63));
    std::mt19937_64 g(5);
    auto make = [&](uint64_t T) {
        std::vector<uint64_t> w(T * L);
        for (uint64_t i = 0; i < w.size(); ++i) w[i] = g() & ((i % L) + 1 == L ? pad : ~0ull);
        for (int k = 0; k < 41; ++k) {
            uint64_t row = g() % T;
            for (uint64_t j = 0; j < L; ++j) w[row * L + j] |= mask[j];
        }
        return w;
    };
    std::vector<uint64_t> wa = make(T1), wb = make(T2), bla(T1 * L), blb(T2 * L);
    for (uint64_t i = 0; i < bla.size(); ++i) bla[i] = ((i % L) + 1 == L && rem) ? rem : 64;
    for (uint64_t i = 0; i < blb.size(); ++i) blb[i] = ((i % L) + 1 == L && rem) ? rem : 64;
    ours::Ciphertext oa(wa.data(), bla.data(), wa.size(), octx), ob(wb.data(), blb.data(), wb.size(), octx);
    ref::Ciphertext ra(wa.data(), bla.data(), wa.size(), rctx), rb(wb.data(), blb.data(), wb.size(), rctx);
    ours::SecretKey osk(octx);
    ref::SecretKey rsk(rctx);
    osk.setKey(key.data(), D);
    rsk.setKey(key.data(), D);

    auto t0 = std::chrono::steady_clock::now();
    ref::Ciphertext rp = ra * rb;
    auto t1 = std::chrono::steady_clock::now();
    int rbit = rsk.decrypt(rp).getValue();
    auto t2 = std::chrono::steady_clock::now();
    ours::Ciphertext op = oa * ob;
    int obit = osk.decrypt(op).getValue();
    auto t3 = std::chrono::steady_clock::now();
    EXPECT(obit == rbit);
    EXPECT(same_ct(op, rp));
    ours::Ciphertext osum = op + oa;
    ref::Ciphertext rsum = rp + ra;
    EXPECT(same_ct(osum, rsum));
    EXPECT(osk.decrypt(osum).getValue() == rsk.decrypt(rsum).getValue());
    std::cout << "config2 N=" << N << " " << T1 << "x" << T2 << " -> " << T1 * T2 << " blocks: bit " << obit
              << "; reference mul " << std::chrono::duration<double, std::milli>(t1 - t0).count() << " ms, decrypt "
              << std::chrono::duration<double, std::milli>(t2 - t1).count() << " ms; drop-in mul+decrypt "
              << std::chrono::duration<double, std::milli>(t3 - t2).count() << " ms" << std::endl;
}

int main() {
    ours::Library::initializeLibrary();
    ref::Library::initializeLibrary();
    encrypted_circuits(1247, 16, 101, 5, 4);
    encrypted_circuits(1247, 16, 102, 1, 1);
    encrypted_circuits(16383, 64, 103, 3, 2);
    encrypted_circuits(65, 2, 104, 6, 5);
    encrypted_circuits(191, 5, 105, 4, 3);  // odd words per block
    encrypted_circuits(63, 4, 106, 3, 3);   // one word per block
    ours::Library::setLazyProducts(true);    // lazy products must be observationally identical
    encrypted_circuits(1247, 16, 107, 4, 3);
    encrypted_circuits(191, 5, 108, 3, 2);
    ours::Library::setLazyProducts(false);
    config2_full_size(1247, 16, 1000, 1000);
    config2_full_size(16383, 64, 300, 300);
    if (failures) {
        std::cerr << failures << " mismatch(es) against the reference" << std::endl;
        return 1;
    }
    std::cout << "diff_vs_reference: identical to the reference on every observable" << std::endl;
    return 0;
}
