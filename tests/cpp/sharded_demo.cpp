// sharded_demo.cpp -- the certFHE C++ API over several GPUs, one process per GPU, no Python and no
// communication library: a shell loop (or any launcher) starts WORLD_SIZE copies with RANK / LOCAL_RANK /
// WORLD_SIZE / CSGN_RENDEZVOUS_DIR set; Library::initializeLibrary() joins them (mailbox handles through the
// directory, peers mapped over NVLink).  Every rank builds the same seeded operands, keeps its block range of
// the left one (Ciphertext::shard), multiplies shard-locally and decrypts: the fold kernel exchanges the counts,
// every rank gets the GLOBAL bit.  Checked against the unsharded computation on the same rank.
// tests/test_cpp_dropin.py runs it with 1 process on a one-GPU box and min(4, n) processes otherwise.
#include "certFHE.h"

#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace certFHE;

static int fails = 0;
#define EXPECT(cond)                                                                  \
    do {                                                                              \
        if (!(cond)) {                                                                \
            ++fails;                                                                  \
            std::fprintf(stderr, "rank %d: FAILED %s (line %d)\n", Library::getRank(), #cond, __LINE__); \
        }                                                                             \
    } while (0)

static Ciphertext sum_of_encryptions(SecretKey &key, const std::vector<int> &bits) {
    Ciphertext acc;
    for (size_t i = 0; i < bits.size(); ++i) {
        Plaintext p(bits[i]);
        Ciphertext c = key.encrypt(p);
        if (i == 0) acc = c; else acc += c;
    }
    return acc;
}

int main() {
    Library::initializeLibrary();
    const int rank = Library::getRank(), world = Library::getWorldSize();
    Context ctx(1247, 16);
    SecretKey key(ctx);
    srand(12345);                                    // the same stream on every rank from here on
    uint64_t pos[16];
    for (int i = 0; i < 16; ++i) pos[i] = 7 + 77 * i;
    key.setKey(pos, 16);

    for (int trial = 0; trial < 6; ++trial) {
        std::vector<int> ba, bb, bd;
        for (int i = 0; i < 37 + 11 * trial; ++i) ba.push_back(rand() % 2);
        for (int i = 0; i < 23 + trial; ++i) bb.push_back(rand() % 2);
        for (int i = 0; i < 5; ++i) bd.push_back(rand() % 2);
        Ciphertext a = sum_of_encryptions(key, ba), b = sum_of_encryptions(key, bb), d = sum_of_encryptions(key, bd);

        // what one process computes
        Ciphertext whole = (a * b) * d;
        const int want_ab = key.decrypt(whole).getValue();
        Ciphertext sum_whole = a + b;
        const int want_sum = key.decrypt(sum_whole).getValue();

        // the same over `world` GPUs
        Ciphertext as = a.shard();
        EXPECT(as.isSharded());
        uint64_t mine = as.getBlocks(), base = a.getBlocks() / world, extra = a.getBlocks() % world;
        EXPECT(mine == base + ((uint64_t)rank < extra ? 1 : 0));
        Ciphertext chain = (as * b) * d;             // shard-local: b and d are replicated
        EXPECT(chain.isSharded());
        EXPECT(chain.getBlocks() == mine * b.getBlocks() * d.getBlocks());
        EXPECT(key.decrypt(chain).getValue() == want_ab);
        Ciphertext sum_sharded = as + b.shard();
        EXPECT(key.decrypt(sum_sharded).getValue() == want_sum);
        // a permuted shard under the permuted key
        Permutation perm(ctx);
        SecretKey pkey = key.applyPermutation(perm);
        Ciphertext pchain = chain.applyPermutation(perm);
        EXPECT(pkey.decrypt(pchain).getValue() == want_ab);
        // misuse is refused, not mis-counted
        bool threw = false;
        try { Ciphertext bad = as + b; (void)bad; } catch (const Error &) { threw = true; }
        EXPECT(threw || world == 1 || true);          // (with one rank a replicated operand is harmless, still refused)
        EXPECT(threw);
        threw = false;
        try { Ciphertext bad = a * as; (void)bad; } catch (const Error &) { threw = true; }
        EXPECT(threw);
    }
    Library::synchronize();
    if (fails == 0) std::printf("sharded_demo rank %d/%d: all checks passed\n", rank, world);
    return fails ? 1 : 0;
}
