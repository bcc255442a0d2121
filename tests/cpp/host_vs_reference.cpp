// host_vs_reference.cpp -- the HOST-ONLY half of the certFHE drop-in against the unmodified
// reference, in one process, no GPU needed: Context, Plaintext, Permutation (generation in the
// reference's rand() order, inverse, composition, soft failures), SecretKey (setKey, size,
// applyPermutation, printing, copy/assign), Helper, Timer.  Ciphertext evaluation is NOT touched
// here -- that needs the engine and is covered by diff_vs_reference.cpp on the GPU box.
#include "certFHE.h"  // ours

#define certFHE certFHE_ref
#include CSGN_REFERENCE_HEADER
#undef certFHE

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

template <class T>
static std::string text(const T &x) {
    std::ostringstream os;
    os << x;
    return os.str();
}

int main() {
    const uint64_t params[][2] = {{1247, 16}, {16383, 64}, {65, 2}, {191, 5}, {63, 4}, {7, 1}};
    for (const auto &nd : params) {
        const uint64_t N = nd[0], D = nd[1];
        ours::Context oc(N, D);
        ref::Context rc(N, D);
        EXPECT(oc.getN() == rc.getN() && oc.getD() == rc.getD() && oc.getS() == rc.getS() &&
               oc.getDefaultN() == rc.getDefaultN());
        EXPECT(text(oc) == text(rc));
        ours::Context oc2(oc);
        oc2.setD(D + 1);
        ref::Context rc2(rc);
        rc2.setD(D + 1);
        EXPECT(oc2.getS() == rc2.getS());

        for (int v = -1; v <= 3; ++v) {
            ours::Plaintext op(v);
            ref::Plaintext rp(v);
            EXPECT(op.getValue() == rp.getValue() && text(op) == text(rp));
        }

        {   // the reference's generation is O(n^2 log n): one seed only at N=16383 (seconds)
            for (unsigned seed = 1; seed <= (N <= 2000 ? 3u : 1u); ++seed) {
                srand(seed);
                ref::Permutation rp(rc);
                srand(seed);
                ours::Permutation op(oc);
                EXPECT(op.getLength() == rp.getLength());
                EXPECT(memcmp(op.getPermutation(), rp.getPermutation(), N * 8) == 0);
                EXPECT(rand() == rand() || true);   // both consumed the same number of draws (checked below)
                srand(seed);
                { ref::Permutation tmp(N); }
                const int after_ref = rand();
                srand(seed);
                { ours::Permutation tmp(N); }
                EXPECT(rand() == after_ref);        // identical rand() consumption
                ours::Permutation oi = op.getInverse();
                ref::Permutation ri = rp.getInverse();
                EXPECT(memcmp(oi.getPermutation(), ri.getPermutation(), N * 8) == 0);
                ours::Permutation oid = op + oi;
                ref::Permutation rid = rp + ri;
                EXPECT(memcmp(oid.getPermutation(), rid.getPermutation(), N * 8) == 0);
                for (uint64_t i = 0; i < N; ++i) EXPECT(oid.getPermutation()[i] == i);
                srand(seed + 100);
                ref::Permutation rq(N);
                srand(seed + 100);
                ours::Permutation oq(N);
                ours::Permutation ocomp = op + oq;
                ref::Permutation rcomp = rp + rq;
                EXPECT(memcmp(ocomp.getPermutation(), rcomp.getPermutation(), N * 8) == 0);
                oq += op;
                rq += rp;
                EXPECT(memcmp(oq.getPermutation(), rq.getPermutation(), N * 8) == 0);
                if (N <= 200) EXPECT(text(op) == text(rp));
                ours::Permutation oshort(3);
                ref::Permutation rshort(3);
                EXPECT((op + oshort).getLength() == (rp + rshort).getLength());   // 0: soft failure
                ours::Permutation okeep(op);
                okeep += oshort;
                EXPECT(memcmp(okeep.getPermutation(), op.getPermutation(), N * 8) == 0);

                // secret key: install positions, permute, print, size, copy, assign
                std::mt19937_64 g(seed);
                std::vector<uint64_t> pool(N);
                for (uint64_t i = 0; i < N; ++i) pool[i] = i;
                for (uint64_t i = 0; i < D; ++i) std::swap(pool[i], pool[i + g() % (N - i)]);
                ours::SecretKey ok(oc);
                ref::SecretKey rk(rc);
                ok.setKey(pool.data(), D);
                rk.setKey(pool.data(), D);
                EXPECT(ok.getLength() == rk.getLength() && ok.size() == rk.size() && text(ok) == text(rk));
                ours::SecretKey okp = ok.applyPermutation(op);
                ref::SecretKey rkp = rk.applyPermutation(rp);
                EXPECT(text(okp) == text(rkp));
                ok.applyPermutation_inplace(oi);
                rk.applyPermutation_inplace(ri);
                EXPECT(text(ok) == text(rk));
                ours::SecretKey ocopy(okp), oassign(oc);
                oassign = okp;
                EXPECT(text(ocopy) == text(okp) && text(oassign) == text(okp));
            }
        }
        // key generation: D distinct positions below N (the reference's are not reproducible)
        ours::SecretKey fresh(oc);
        EXPECT(fresh.getLength() == D);
        for (uint64_t i = 0; i < D; ++i) {
            EXPECT(fresh.getKey()[i] < N);
            for (uint64_t j = 0; j < i; ++j) EXPECT(fresh.getKey()[i] != fresh.getKey()[j]);
        }
    }
    const uint64_t arr[4] = {5, 9, 1, 7};
    EXPECT(ours::Helper::exists(arr, 4, 9) == ref::Helper::exists(arr, 4, 9));
    EXPECT(ours::Helper::exists(arr, 4, 2) == ref::Helper::exists(arr, 4, 2));
    ours::Timer t("t");
    t.start();
    EXPECT(t.stop() >= 0.0 && t.getValue() >= 0.0);
    if (failures) {
        std::cerr << failures << " mismatch(es) against the reference" << std::endl;
        return 1;
    }
    std::cout << "host_vs_reference: host-side classes identical to the reference" << std::endl;
    return 0;
}
