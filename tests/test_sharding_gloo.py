"""CPU, world_size 2 over gloo: the host-side logic of the N>1 path.

The kernels cannot run here, so the oracle stands in for the per-rank compute (a test
double with the engine.Ciphertext surface).  What is under test is everything around it:
csgn_shard_range, the contiguity and i-major order of the sharded product, chained
shard-local multiplies, and the all-reduce that finishes decrypt."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.pyoracle import Oracle, random_blocks, random_key, words_per_block

N, D = 1247, 2


class OracleCiphertext:
    """engine.Ciphertext look-alike computed by the oracle (tests only)."""

    def __init__(self, words, L, o):
        self.w, self.L, self.o = np.ascontiguousarray(words, dtype=np.uint64), L, o

    @property
    def n_blocks(self):
        return self.w.size // self.L

    def __mul__(self, other):
        return OracleCiphertext(self.o.mul(self.w, other.w, self.L), self.L, self.o)

    def applyPermutation(self, perm):
        return OracleCiphertext(self.o.permute_all(self.w, N, perm), self.L, self.o)


class OracleKey:
    def __init__(self, s, o):
        self.s, self.o = s, o

    def count_satisfied(self, ct):
        return self.o.count_satisfied(ct.w, N, self.s)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from csgn_b200 import sharding
        o = Oracle()
        L = words_per_block(N)
        rng = np.random.default_rng(123)                      # same stream on every rank
        a, b, d = random_blocks(rng, 37, N), random_blocks(rng, 11, N), random_blocks(rng, 5, N)
        s = random_key(rng, N, D)
        perm = rng.permutation(N).astype(np.uint64)

        class Ctx:
            pass
        ctx = Ctx()
        ctx.L = L
        make = lambda w, c: OracleCiphertext(w, L, o)
        A = sharding.ShardedCiphertext.scatter_from_host(a, ctx, make)
        first, count = sharding.shard_range(37, rank, world)
        assert (A.first, A.count, A.global_blocks) == (first, count, 37)

        # (a*b)*d, shard-local with replicated right operands
        B, Dd = OracleCiphertext(b, L, o), OracleCiphertext(d, L, o)
        P1 = A.mul_replicated(B)
        P2 = P1.mul_replicated(Dd)
        assert (P1.first, P1.count, P1.global_blocks) == (first * 11, count * 11, 37 * 11)
        assert (P2.first, P2.count, P2.global_blocks) == (first * 55, count * 55, 37 * 55)
        full = o.mul(o.mul(a, b, L), d, L)                    # what one process would compute
        assert np.array_equal(P2.local.w, full[P2.first * L:(P2.first + P2.count) * L])

        # decrypt: local fold + one all-reduce == the single-process answer
        key = OracleKey(s, o)
        bit = P2.decrypt(key)
        assert bit == o.decrypt(full, N, s)
        # the same into a caller's tensor (the local count must be written before the all-reduce)
        mine = torch.full((1,), 12345, dtype=torch.int64)
        assert P2.decrypt(key, counts_out=mine) == bit
        assert int(mine.item()) == o.count_satisfied(full, N, s)
        # batched form: P counts, one collective
        counts = torch.tensor([key.count_satisfied(P1.local), key.count_satisfied(P2.local)], dtype=torch.int64)
        sharding.allreduce_counts(counts)
        want = [o.count_satisfied(o.mul(a, b, L), N, s), o.count_satisfied(full, N, s)]
        assert counts.tolist() == want and sharding.parity(counts).tolist() == [w & 1 for w in want]

        # permutation is shard-local and commutes with sharding
        PP = P1.permute(perm)
        full_p = o.permute_all(o.mul(a, b, L), N, perm)
        assert np.array_equal(PP.local.w, full_p[PP.first * L:(PP.first + PP.count) * L])
        k2 = OracleKey(o.key_permute(N, s, perm), o)
        assert PP.decrypt(k2) == P1.decrypt(key)

        # gather the shards and compare with the unsharded product, in rank order
        gathered = [None] * world
        dist.all_gather_object(gathered, P1.local.w)
        assert np.array_equal(np.concatenate(gathered), o.mul(a, b, L))
        results[rank] = "ok"
    except Exception as e:  # surfaces in the parent
        results[rank] = "FAILED: %r" % (e,)
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_mul_decrypt_permute_world2():
    world = 2
    with mp.Manager() as m:
        results = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}


def test_single_process_is_the_same_code_path():
    """world size 1 (no process group): allreduce is the identity, ranges are the whole."""
    from csgn_b200 import sharding
    assert sharding.world() == 1 and sharding.rank() == 0
    assert sharding.shard_range(1000, 0, 1) == (0, 1000)
    c = torch.tensor([5, 8], dtype=torch.int64)
    assert sharding.allreduce_counts(c).tolist() == [5, 8] and sharding.parity(c).tolist() == [1, 0]
