"""Sharded decrypt with the exchange fused into the fold kernel (csrc/peer.cuh, csgn_comm_*).

CPU: the mailbox addressing arithmetic.  GPU: the same kernel path at world size 1 (its own
mailbox is the only one), checked against the oracle; and, where the box has more than one
GPU, tests/peer_worker.py under torchrun on min(4, n) GPUs (the driver's GPU box has one, so that test
runs in `gpurun --gpus 2` sessions; its log is committed under profiles/)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle.pyoracle import random_blocks, random_key, words_per_block

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_mailbox_slot_and_tag_arithmetic():
    from csgn_b200 import engine
    ring = 256
    seen = {}
    for seq in list(range(0, 3 * ring + 5)) + [ring * 0xFFFFFF - 1, ring * 0xFFFFFF, ring * 0xFFFFFF + 1, 2**40 + 17]:
        slot, tag = engine.comm_slot_tag(seq)
        assert slot == seq % ring
        assert 1 <= tag < 2**24                      # never 0: a zero-initialised mailbox word is never valid
        assert tag == (seq // ring) % 0xFFFFFF + 1
        if seq >= ring:                              # consecutive uses of one slot carry different tags
            assert engine.comm_slot_tag(seq - ring)[1] != tag
        seen[seq] = (slot, tag)
    assert engine.COMM_MAX_PENDING < ring // 2       # a fast rank cannot lap a slot a slow rank still has to read


def _planted(rng, T, N, s, k):
    w = random_blocks(rng, T, N).reshape(T, -1)
    mask = np.zeros(w.shape[1], dtype=np.uint64)
    for p in s:
        mask[int(p) >> 6] |= np.uint64(1 << (63 - (int(p) & 63)))
    if T:
        w[rng.choice(T, size=min(T, k), replace=False)] |= mask
    return w.reshape(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("N,D", [(1247, 16), (16383, 64), (191, 5), (2048, 8), (4097, 3)])
def test_fused_exchange_world_size_1(engine, oracle, N, D):
    import torch
    comm = engine.PeerComm(0, 1)
    ctx = engine.Context(N, D)
    rng = np.random.default_rng([N, 3])
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    sizes = (1, 33, 1000, 20011)
    words = [_planted(rng, T, N, s, 1 + T // 5) for T in sizes]
    cts = [engine.Ciphertext.from_host(w, ctx) for w in words]
    want = [oracle.count_satisfied(w, N, s) for w in words]
    assert all(w > 0 for w in want)
    for ct, wn in zip(cts, want):
        assert comm.decrypt(key, ct) == (wn & 1, wn)
    n = len(cts)
    totals = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    local = torch.full((n,), -1, dtype=torch.int64, device="cuda")
    before = engine.launch_count()
    for i, ct in enumerate(cts):
        comm.push(key, ct, collect_n=n if i == n - 1 else 0, device_totals_ptr=totals.data_ptr(),
                  device_local_ptr=local.data_ptr() + 8 * i)
    engine.sync()
    assert engine.launch_count() == before + n       # fold, push and collect: one launch per ciphertext
    assert totals.tolist() == want and local.tolist() == want
    totals.fill_(-1)
    comm.push_batch(key, cts, totals.data_ptr())     # the same through the batch call (folds spread over the lanes)
    engine.sync()
    assert totals.tolist() == want and comm.pending == 0
    for ct in cts:
        comm.push(key, ct)
    assert comm.pending == n
    totals.fill_(-1)
    comm.collect(2, totals.data_ptr())               # the most recent two
    engine.sync()
    assert totals.tolist()[:2] == want[-2:] and comm.pending == 0


@pytest.mark.gpu
def test_fused_exchange_ring_wrap_and_limits(engine, oracle):
    import torch
    N, D = 1247, 16
    comm = engine.PeerComm(0, 1)
    ctx = engine.Context(N, D)
    rng = np.random.default_rng(11)
    s = random_key(rng, N, D)
    key = engine.SecretKey(ctx, s)
    words = [_planted(rng, T, N, s, 3 + i) for i, T in enumerate((70, 90, 110))]
    cts = [engine.Ciphertext.from_host(w, ctx) for w in words]
    want = [oracle.count_satisfied(w, N, s) for w in words]
    rounds = 200                                     # 600 pushes: the 256-slot ring wraps twice
    totals = torch.zeros((rounds, 3), dtype=torch.int64, device="cuda")
    for r in range(rounds):
        for i, ct in enumerate(cts):
            comm.push(key, ct, collect_n=3 if i == 2 else 0, device_totals_ptr=totals[r].data_ptr())
    engine.sync()
    assert totals.tolist() == [want] * rounds
    # a lagged collect: the launch closing batch k returns batch k-1; a final collect fetches the last
    fresh = engine.PeerComm(0, 1)
    with pytest.raises(engine.CsgnError, match="only"):
        fresh.push(key, cts[0], collect_n=1, device_totals_ptr=totals.data_ptr(), lag=1)   # nothing to trail yet
    got = torch.full((4, 3), -1, dtype=torch.int64, device="cuda")
    order = [[0, 1, 2], [2, 0, 1], [1, 1, 0]]
    for b, idx in enumerate(order):
        for i, k in enumerate(idx):
            closing = i == 2
            if closing and b == 0:
                fresh.push(key, cts[k], collect_n=3, device_totals_ptr=got[0].data_ptr())
            elif closing:
                fresh.push(key, cts[k], collect_n=3, device_totals_ptr=got[b].data_ptr(), lag=3)
            else:
                fresh.push(key, cts[k])
    fresh.collect(3, got[3].data_ptr())
    engine.sync()
    w = lambda idx: [want[k] for k in idx]
    assert got.tolist() == [w(order[0]), w(order[0]), w(order[1]), w(order[2])]
    # an empty shard still pushes its zero
    empty = engine.Ciphertext.empty(0, ctx)
    assert comm.decrypt(key, empty) == (0, 0)
    # limits
    for _ in range(engine.COMM_MAX_PENDING):
        comm.push(key, cts[0])
    with pytest.raises(engine.CsgnError, match="without a collect"):
        comm.push(key, cts[0])
    with pytest.raises(engine.CsgnError, match="exceeds"):
        comm.collect(engine.COMM_MAX_PENDING + 1, totals.data_ptr())
    with pytest.raises(engine.CsgnError, match="exceeds"):
        comm.collect(engine.COMM_MAX_PENDING, totals.data_ptr(), lag=1)
    comm.collect(engine.COMM_MAX_PENDING, totals.data_ptr())
    engine.sync()
    assert totals.reshape(-1)[:engine.COMM_MAX_PENDING].tolist() == [want[0]] * engine.COMM_MAX_PENDING
    with pytest.raises(engine.CsgnError, match="words per block"):
        comm.push(engine.SecretKey(engine.Context(191, 2), np.array([1, 5], dtype=np.uint64)), cts[0])
    with pytest.raises(ValueError):
        engine.PeerComm(0, 2)                        # world > 1 needs the handle exchange


@pytest.mark.gpu
def test_fused_exchange_across_gpus(engine):
    import torch
    n = min(4, torch.cuda.device_count())
    if n < 2:
        pytest.skip("one GPU visible: the cross-GPU run is tests/peer_worker.py under `gpurun --gpus 2`")
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "peer_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("groups of checks passed") == n, r.stdout[-4000:]


# ---------------------------------------------------------------------------
# the mailbox protocol as a model (CPU): ranks drift freely between collects, yet no slot is ever
# overwritten before every rank that still has to read it has done so
# ---------------------------------------------------------------------------
def _simulate(world, program, rng, ring=256, max_steps=10**6):
    """program: list of ("push",) / ("close", n, lag) executed by EVERY rank (the collective contract).
    A close publishes the rank's unpublished pushes to every mailbox, then blocks until the window
    [last-lag-n+1, last-lag] of every rank has arrived.  The scheduler picks a runnable rank at random.
    Returns the number of stale reads (a collected word that is not the push it should be)."""
    tag_of = lambda s_: (s_ // ring) % 0xFFFFFF + 1          # peer.cuh's tag (checked against the C function above)
    box = [[[None] * world for _ in range(ring)] for _ in range(world)]       # box[q][slot][r] = (tag, seq)
    pc = [0] * world                 # program counter per rank
    seq = [0] * world                # next push
    published = [0] * world
    waiting = [None] * world         # (first, last) window a rank blocks on
    stale = 0
    for _ in range(max_steps):
        runnable = []
        for r in range(world):
            if waiting[r] is not None:
                first, last = waiting[r]
                ok = all(box[r][s % ring][q] is not None and box[r][s % ring][q][0] == tag_of(s)
                         for s in range(first, last + 1) for q in range(world))
                if ok:
                    runnable.append(r)
            elif pc[r] < len(program):
                runnable.append(r)
        if not runnable:
            break
        r = runnable[int(rng.integers(len(runnable)))]
        if waiting[r] is not None:                       # the collect completes: read the window
            first, last = waiting[r]
            for s in range(first, last + 1):
                for q in range(world):
                    if box[r][s % ring][q][1] != s:
                        stale += 1
            waiting[r] = None
            pc[r] += 1
            continue
        op = program[pc[r]]
        if op[0] == "push":
            seq[r] += 1
            pc[r] += 1
        else:
            _, n, lag = op
            for s in range(published[r], seq[r]):        # publish: one word per (push, rank)
                for q in range(world):
                    box[q][s % ring][r] = (tag_of(s), s)
            published[r] = seq[r]
            last = seq[r] - 1 - lag
            waiting[r] = (last - n + 1, last)
    assert all(p == len(program) for p in pc), "deadlock in the model"
    return stale


def _random_program(rng, batches, max_pending, lagged):
    prog, prev = [], 0
    for _ in range(batches):
        n = int(rng.integers(1, max_pending // 2 + 1))
        prog += [("push",)] * n
        if lagged and prev and prev + n <= max_pending:
            prog.append(("close", prev, n))              # the previous batch: never waits for a slow peer
        else:
            prog.append(("close", n, 0))
        prev = n
    return prog


def test_mailbox_protocol_model_never_reads_a_lapped_slot():
    from csgn_b200 import engine
    rng = np.random.default_rng(2025)
    for world in (2, 3, 8):
        for lagged in (False, True):
            prog = _random_program(rng, 60, engine.COMM_MAX_PENDING, lagged)     # ~1000 pushes: the ring wraps 4 times
            assert _simulate(world, prog, rng) == 0
    class Greedy:                                        # always pick the lowest runnable rank
        def integers(self, n):
            return 0
    # the model does catch a protocol that breaks the bound: batches of 6 on a ring of 8 (bound: fewer than 4).  Rank 0
    # finishes its first collect and publishes batch 2 over slots rank 1 has not read yet; rank 1 then never sees the
    # tags it waits for (on the GPU: the collect would time out)
    with pytest.raises(AssertionError, match="deadlock"):
        _simulate(2, ([("push",)] * 6 + [("close", 6, 0)]) * 3, Greedy(), ring=8)
    assert _simulate(2, ([("push",)] * 3 + [("close", 3, 0)]) * 6, Greedy(), ring=8) == 0      # within the bound: clean
    # deterministic adversary: rank 0 runs ahead as far as the contract allows, rank 1 lags -- still clean
    prog = _random_program(rng, 40, engine.COMM_MAX_PENDING, True)
    assert _simulate(4, prog, Greedy()) == 0
