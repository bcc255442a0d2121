"""One rank of the multi-GPU check of the fused decrypt + exchange kernel (csrc/peer.cuh).

Launched by tests/test_peer_exchange.py (and by hand on a multi-GPU box) as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/peer_worker.py
Every rank holds the same seeded global ciphertexts, uploads only its csgn_shard_range of
each, and checks the exchanged totals against the oracle's count over the WHOLE ciphertext.
Exit code 0 = every check passed on this rank.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    os.environ.setdefault("CSGN_PEER_TIMEOUT_MS", "5000")
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from csgn_b200 import engine as eng, sharding
    from oracle.pyoracle import Oracle, random_blocks, random_key, words_per_block
    eng.init(local)
    o = Oracle()
    comm = sharding.connect_peers()
    assert (comm.rank, comm.world) == (rank, world)

    def planted(rng, T, N, s, k):
        w = random_blocks(rng, T, N).reshape(T, -1)
        mask = np.zeros(w.shape[1], dtype=np.uint64)
        for p in s:
            mask[int(p) >> 6] |= np.uint64(1 << (63 - (int(p) & 63)))
        w[rng.choice(T, size=min(T, k), replace=False)] |= mask
        return w.reshape(-1)

    checked = 0
    for N, D, sizes in ((1247, 16, (1, 5, 1000, 40001)), (16383, 64, (3, 700)), (191, 5, (257,)), (2048, 8, (999,)), (4097, 7, (131, 4001)), (3197, 4, (77,))):
        L = words_per_block(N)
        ctx = eng.Context(N, D)
        rng = np.random.default_rng([N, 77])                 # the same stream on every rank
        s = random_key(rng, N, D)
        key = eng.SecretKey(ctx, s)
        globals_, shards, want = [], [], []
        for T in sizes:
            w = planted(rng, T, N, s, 1 + T // 7)
            first, count = eng.shard_range(T, rank, world)   # may be empty (T < world)
            shards.append(eng.Ciphertext.from_host(w[first * L:(first + count) * L], ctx))
            want.append(o.count_satisfied(w, N, s))
            globals_.append(w)
        # (1) blocking sharded decrypt, one ciphertext at a time
        for sh, wn in zip(shards, want):
            bit, total = comm.decrypt(key, sh)
            assert (bit, total) == (wn & 1, wn), (N, rank, bit, total, wn)
            checked += 1
        # (2) a batch: n-1 pushes, the last launch collects all n in the same kernel
        n = len(shards)
        totals = torch.full((n,), -1, dtype=torch.int64, device=dev)
        locals_ = torch.full((n,), -1, dtype=torch.int64, device=dev)
        for i, sh in enumerate(shards):
            comm.push(key, sh, collect_n=n if i == n - 1 else 0, device_totals_ptr=totals.data_ptr(),
                      device_local_ptr=locals_.data_ptr() + 8 * i)
        eng.sync()
        assert totals.tolist() == want, (N, rank, totals.tolist(), want)
        mine = [o.count_satisfied(g[eng.shard_range(T, rank, world)[0] * L:
                                    (eng.shard_range(T, rank, world)[0] + eng.shard_range(T, rank, world)[1]) * L], N, s)
                for g, T in zip(globals_, sizes)]
        assert locals_.tolist() == mine
        # (3) pushes now, a stand-alone collect later
        for sh in shards:
            comm.push(key, sh)
        assert comm.pending == n
        totals.fill_(-1)
        comm.collect(n, totals.data_ptr())
        eng.sync()
        assert totals.tolist() == want and comm.pending == 0
        checked += 2
        # the batch call: folds spread over the library's lanes, one closing launch
        totals.fill_(-1)
        comm.push_batch(key, shards, totals.data_ptr())
        eng.sync()
        assert totals.tolist() == want
        # agrees with the all-reduce it replaces
        ar = locals_.clone()
        dist.all_reduce(ar)
        assert ar.tolist() == want

    # (4) many rounds: the mailbox ring (256 slots) wraps several times, ranks drift freely between collects
    N, D = 1247, 16
    L = words_per_block(N)
    ctx = eng.Context(N, D)
    rng = np.random.default_rng(5)
    s = random_key(rng, N, D)
    key = eng.SecretKey(ctx, s)
    w = planted(rng, 64 * world + 3, N, s, 40)
    first, count = eng.shard_range(64 * world + 3, rank, world)
    sh = eng.Ciphertext.from_host(w[first * L:(first + count) * L], ctx)
    want1 = o.count_satisfied(w, N, s)
    rounds, per = 120, 7
    totals = torch.zeros((rounds, per), dtype=torch.int64, device=dev)
    for r in range(rounds):
        for i in range(per):
            comm.push(key, sh, collect_n=per if i == per - 1 else 0, device_totals_ptr=totals[r].data_ptr())
        if r % 16 == rank % 16:
            eng.sync()                                        # let the ranks fall out of step
    eng.sync()
    assert torch.all(totals == want1), (rank, totals[totals != want1][:8].tolist(), want1)
    # the same with lagged collects: the launch closing round r returns round r-1 (no waiting on peers)
    lag_totals = torch.zeros((rounds, per), dtype=torch.int64, device=dev)
    shards2 = [sh, eng.Ciphertext.from_host(w[first * L:(first + count // 2) * L], ctx)]
    want2 = [want1, None]
    half = torch.tensor([key.count_satisfied(shards2[1])], dtype=torch.int64, device=dev)
    dist.all_reduce(half)
    want2[1] = int(half.item())
    for r in range(rounds):
        for i in range(per):
            ct = shards2[(r + i) % 2]
            if i < per - 1:
                comm.push(key, ct)
            elif r == 0:
                comm.push(key, ct, collect_n=per, device_totals_ptr=lag_totals[0].data_ptr())
            else:
                comm.push(key, ct, collect_n=per, device_totals_ptr=lag_totals[r - 1].data_ptr(), lag=per)
        if r % 16 == (rank + 3) % 16:
            eng.sync()
    comm.collect(per, lag_totals[rounds - 1].data_ptr())
    eng.sync()
    expect = [[want2[(r + i) % 2] for i in range(per)] for r in range(rounds)]
    assert lag_totals.tolist() == expect, (rank, want2)
    checked += 2

    # (5) a rank that never arrives: the collect times out, the call fails, the GPU is not left spinning
    dist.barrier()
    if world > 1:
        os.environ["CSGN_PEER_TIMEOUT_MS"] = "300"
        short = sharding.connect_peers()                      # a second communicator with the short timeout
        if rank == 0:
            try:
                short.decrypt(key, sh)
                raise AssertionError("expected a timeout")
            except eng.CsgnError as e:
                assert e.code == -7, e
        dist.barrier()
        del short
        checked += 1
    eng.sync()
    dist.barrier()
    print("peer_worker rank %d/%d: %d groups of checks passed" % (rank, world, checked), flush=True)
    del comm
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
