#!/usr/bin/env python
"""bench.py -- the headline measurement of BASELINE.json on B200.

Metric: output blocks/s of the ciphertext hot path -- multiply (all-pairs AND of
T1 x T2 blocks) followed by decrypt of the product -- at Context(1247,16),
1000 x 1000 -> 1,000,000 output blocks per ciphertext pair (BASELINE.json configs[1]).

A step is one pass over a batch of `pairs` independent ciphertext pairs:
    [multiply pair 0 .. P-1]  then  [decrypt product 0 .. P-1]
so every product (160 MB) is written and, P-1 products later, read back: the batch
(P x 160 MB) is far larger than the 126 MB L2 and both kernels run against HBM.

  value     device-resident operands and outputs, no host traffic in the timed region
  e2e       the same batch through the public C ABI from pinned HOST operands:
            csgn_buf_upload x2 -> csgn_mul -> csgn_decrypt_count_async, one D2H of
            the P counts per step
  roofline  multiply kernel: 160 B written per output block / CUDA-event time of the
            multiply phase, against the measured HBM figure of MEASURED_PEAKS.json
  cpu_baseline  the unmodified reference (oracle/_ref) or the oracle port, 1 thread

N > 1 (torchrun, one process per GPU): weak scaling.  The left operand of every pair
has 1000*N blocks and is sharded by contiguous block range (csgn_shard_range); the
right operand is replicated; every rank multiplies and folds its own 1M-block shard
and the per-pair satisfied-block counts are summed by ONE NCCL all-reduce per step
(P words) -- decrypt is the parity of the sum.

`--impl reference` times the reference's own CPU implementation (public operator*
and SecretKey::decrypt) on the host cores, one replica per core.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (N, D, T1, T2, description)
    "cfg2": (1247, 16, 1000, 1000, "Context(1247,16): 1000x1000 -> 1M output blocks, multiply then decrypt"),
    "cfg5": (16383, 64, 300, 300, "Context(16383,64): 300x300 -> 90k output blocks, multiply then decrypt"),
    # BASELINE.json configs[3]: deep product chain (a*b)*d; per GPU 1000 x 1000 x chain_d blocks, the left
    # operand sharded by block range, b and d replicated, decrypt finished by a one-word all-reduce.
    # chain_d = 125 gives 1.25e8 blocks = 20 GB per GPU, 1e9 blocks at 8 GPUs.
    "cfg4": (1247, 16, 1000, 1000, "Context(1247,16): chain (a*b)*d, 1000 x 1000 x chain_d blocks per GPU, decrypt + all-reduce"),
}
METRIC = "ctxt-mul+decrypt output blocks/s"
UNIT = "blocks/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=16, help="independent ciphertext pairs per step")
    ap.add_argument("--chain-d", type=int, default=125, help="cfg4: blocks of the third operand (125 -> 20 GB/GPU)")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: decrypt's cross-GPU sum inside the fold kernel over NVLink mailboxes (peer) "
                         "or as a separate NCCL all-reduce (nccl, the baseline it replaces)")
    ap.add_argument("--collect", default="lagged", choices=["lagged", "same-step"],
                    help="peer exchange: the launch closing step k collects step k-1's sums (never waits for a slower "
                         "rank; the last step's sums are collected before the timed region ends) or its own step's")
    ap.add_argument("--enqueue", default="batch", choices=["batch", "streams"],
                    help="batch: one csgn_mul_into_batch + one csgn_decrypt_count_batch_async per step -- the library "
                         "spreads the independent pairs over its internal lanes (default); streams: one call per pair, "
                         "bench.py itself alternates --streams CUDA streams (what the batch calls do inside)")
    ap.add_argument("--e2e-enqueue", default="batch", choices=["batch", "pairs"],
                    help="e2e with --enqueue batch: batch = upload all pairs, csgn_mul_batch, one batched fold; pairs = "
                         "upload/multiply/fold/free pair by pair on alternating streams (each product is folded while "
                         "part of it is still in L2)")
    ap.add_argument("--streams", type=int, default=2,
                    help="enqueue the independent pairs of a step round-robin on this many CUDA streams (csgn_set_stream "
                         "between calls): the tail of one kernel overlaps the ramp of the next pair's")
    ap.add_argument("--t1", type=int, default=0, help="blocks of the left operand (overrides the workload's; with --scaling "
                                                      "strong: of the WHOLE left operand, sharded over the ranks)")
    ap.add_argument("--t2", type=int, default=0, help="blocks of the right operand (overrides the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = every rank multiplies t1 x t2 (default, the headline); strong = the t1 blocks of the "
                         "left operand are split over the ranks (SURVEY 8d, the cfg5 sweep)")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the side measurements of permute and add (GPU and reference CPU) reported under other_kernels")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def words_per_block(N):
    return N // 64 + (1 if N % 64 else 0)


def seeded_blocks(rng, T, N):
    L = words_per_block(N)
    w = rng.integers(0, 2**64, size=(T, L), dtype=np.uint64)
    rem = N % 64
    if rem:
        w[:, L - 1] &= np.uint64((0xFFFFFFFFFFFFFFFF << (64 - rem)) & 0xFFFFFFFFFFFFFFFF)
    return w.reshape(-1)


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel_prefix, workload):
    """dram bytes (read + write) per launch of the dominant kernel, from the committed ncu --set full
    capture of this workload (profiles/r1b_<workload>_ncu_summary.json), or None."""
    for tag in ("r1c", "r1b", "r1"):                      # the latest capture first
        try:
            with open(os.path.join(ROOT, "profiles", "%s_%s_ncu_summary.json" % (tag, workload))) as f:
                for name, k in json.load(f)["kernels"].items():
                    if name.startswith(kernel_prefix):
                        return k["dram_traffic_bytes_per_launch"]
        except Exception:
            pass
    return None


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons during the timed region (NVML, ~2 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = str(e)

    NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
             0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
             0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                mhz = self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                self.samples.append((time.perf_counter(), mhz, r))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and the throttle reasons of the samples taken inside [t_begin, t_end] (perf_counter): the
        thread is started a little before the timed region so that NVML is warm, and only what falls inside counts."""
        self._stop_evt.set()
        self.join(timeout=2)
        inside = [x for x in self.samples if (t_begin is None or x[0] >= t_begin) and (t_end is None or x[0] <= t_end)]
        note = None
        if not inside and self.samples and t_begin is not None:
            # a timed region shorter than one NVML round trip (a few steps only): take the samples closest to it --
            # the GPU ran the same kernels (warm-up, e2e) right before and after
            mid = 0.5 * (t_begin + (t_end if t_end is not None else t_begin))
            inside = sorted(self.samples, key=lambda x: abs(x[0] - mid))[:3]
            note = "timed region shorter than the sampling period: the %d samples nearest to it" % len(inside)
        reasons = set()
        for _, _, r in inside:
            for bit, name in self.NAMES.items():
                if r & bit and name != "gpu_idle":
                    reasons.add(name)
        med = float(np.median([x[1] for x in inside])) if inside else None
        out = {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(reasons), "samples": len(inside)}
        if note:
            out["note"] = note
        return out


# ---------------------------------------------------------------------------
# CPU arms (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------
def cpu_reference_once(N, D, T1, T2, threads, reps, opt="O3"):
    """(kind, mul_s, dec_s) for `threads` replicas of a T1 x T2 multiply + decrypt."""
    from oracle import pyoracle
    if pyoracle.ref_available(opt):
        ref = pyoracle.Ref(opt)
        m, d, _ = ref.bench_mul_decrypt(N, D, T1, T2, threads=threads, reps=reps, seed=1)
        return "reference", m, d
    # the reference build did not travel: time the C restatement (one thread)
    o = pyoracle.Oracle()
    rng = np.random.default_rng(1)
    a, b = seeded_blocks(rng, T1, N), seeded_blocks(rng, T2, N)
    s = rng.permutation(N)[:D].astype(np.uint64)
    best_m = best_d = 1e300
    for _ in range(reps):
        t0 = time.perf_counter()
        prod = o.mul(a, b, words_per_block(N))
        t1 = time.perf_counter()
        o.decrypt(prod, N, s)
        t2 = time.perf_counter()
        best_m, best_d = min(best_m, t1 - t0), min(best_d, t2 - t1)
    return "port", best_m, best_d


def cpu_baseline(N, D, T1, T2):
    kind, m, d = cpu_reference_once(N, D, T1, T2, threads=1, reps=3)
    blocks = T1 * T2
    out = {"value": blocks / (m + d), "unit": UNIT, "cores": 1, "kind": kind,
           "sample": "one %dx%d pair (%d output blocks): public operator* %.3f s + SecretKey::decrypt %.3f s, "
                     "best of 3, single thread (the reference has no threading), built -O3 -DNDEBUG"
                     % (T1, T2, blocks, m, d),
           "mul_blocks_per_s": blocks / m, "decrypt_blocks_per_s": blocks / d,
           "host_cores_available": os.cpu_count()}
    from oracle import pyoracle
    if kind == "reference" and pyoracle.ref_available("O0"):
        # what the shipped CMakeLists (no build type, hence no -O flag) really produces; a quarter-size sample
        _, m0, d0 = cpu_reference_once(N, D, max(1, T1 // 4), T2, threads=1, reps=1, opt="O0")
        out["as_shipped_no_opt_flag_blocks_per_s"] = max(1, T1 // 4) * T2 / (m0 + d0)
    return out


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, D, T1, T2, desc = WORKLOADS[args.workload]
    threads = max(1, os.cpu_count() or 1)
    # bounded sample: a quarter of the left operand per replica keeps one step near 1 s
    T1s = max(1, T1 // 4)
    try:
        import psutil
        per_replica = T1s * T2 * (words_per_block(N) * 8 * 4 + N) * 1.2   # v+bitlen twice + unpack scratch
        threads = max(1, min(threads, int(psutil.virtual_memory().available * 0.6 / per_replica)))
    except Exception:
        pass
    times = []
    kind = "reference"
    for i in range(args.warmup + args.steps):
        kind, m, d = cpu_reference_once(N, D, T1s, T2, threads=threads, reps=1)
        if i >= args.warmup:
            times.append(m + d)
    blocks = threads * T1s * T2
    total = float(np.sum(times))
    value = blocks * len(times) / total
    sample = ("%d replica threads x one %dx%d pair per step (%d output blocks/step), public operator* + "
              "SecretKey::decrypt; harness-level parallelism, the reference itself is single-threaded"
              % (threads, T1s, T2, blocks))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "sample": sample},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def other_kernels(eng, torch, ctx, vo, N, D, L, peak, with_cpu):
    """Permute and add on products the bench already holds (reported beside the headline, each
    against the HBM roofline at 16*L bytes per block, with the unmodified reference's public calls on a bounded
    sample).  Never part of `value`; any failure here is reported, not raised."""
    out = {}
    try:
        T = vo[0].n_blocks
        nb = len(vo) // 2 * 2
        if nb < 2:
            return {"skipped": "needs two products to ping-pong between (--pairs >= 2)"}
        perm_np = np.random.default_rng(3).permutation(N).astype(np.uint64)
        perm = eng.Permutation(ctx, perm_np)

        def timed(fn, reps):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3 / reps

        def permute_pass():                      # every even product into its odd neighbour: nothing stays in L2
            for i in range(0, nb, 2):
                vo[i].permute_into(perm, vo[i + 1])
        t = timed(permute_pass, 3) / (nb // 2)
        gbs = T * 16 * L / t / 1e9
        out["permute"] = {"blocks_per_s": T / t, "gbs_read_plus_write": gbs, "frac_of_peak": gbs / peak,
                          "avg_launch_us": t * 1e6, "blocks_per_launch": T}

        def add_pass():
            for i in range(0, nb, 2):
                s_ = vo[i] + vo[i + 1]           # csgn_concat into a fresh buffer
                del s_
        t = timed(add_pass, 3) / (nb // 2)
        gbs = 2 * T * 16 * L / t / 1e9
        out["add"] = {"blocks_per_s": 2 * T / t, "gbs_read_plus_write": gbs, "frac_of_peak": gbs / peak,
                      "avg_launch_us": t * 1e6, "blocks_per_launch": 2 * T}
        if with_cpu:
            from oracle import pyoracle
            if pyoracle.ref_available():
                ref = pyoracle.Ref()
                rng = np.random.default_rng(5)
                Tp = 20000 if N < 4096 else 1500          # the reference unpacks 8 bytes per BIT: 43 us per block at N=1247
                v = seeded_blocks(rng, Tp, N)
                h = ref.ct(v, N, D)
                t0 = time.perf_counter()
                hp = ref.lib.ref_ct_permute(h, perm_np.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), N)
                tp = time.perf_counter() - t0
                ref.ct_free(hp)
                Ta = 100000 if N < 4096 else 8000
                h2 = ref.ct(seeded_blocks(rng, Ta, N), N, D)
                t0 = time.perf_counter()
                hs = ref.lib.ref_ct_add(h2, h)
                ta = time.perf_counter() - t0
                for x in (h, h2, hs):
                    ref.ct_free(x)
                out["permute"]["cpu_reference_blocks_per_s"] = Tp / tp
                out["permute"]["cpu_sample"] = ("public applyPermutation on a %d-block ciphertext, %.2f s, one thread; the "
                                                "reference walks every block but RETURNS only block 0 (src/Ciphertext.cpp:33-40)"
                                                % (Tp, tp))
                out["add"]["cpu_reference_blocks_per_s"] = (Ta + Tp) / ta
                out["add"]["cpu_sample"] = "public operator+ of %d + %d blocks, %.3f s, one thread" % (Ta, Tp, ta)
    except Exception as e:  # noqa: BLE001
        out["error"] = "%s: %s" % (type(e).__name__, e)
    return out



def connect_exchange(args, world, dev):
    """(comm, description).  comm is None at N=1 and for --exchange nccl.  Every rank must take the same
    path, so a rank that cannot map its peers' mailboxes makes all of them fall back to NCCL -- loudly."""
    if world == 1:
        return None, "single GPU: no exchange"
    if args.exchange == "nccl":
        return None, "separate NCCL all-reduce of the counts"
    import torch
    import torch.distributed as dist
    from csgn_b200 import sharding
    comm, err = None, ""
    try:
        comm = sharding.connect_peers()
    except Exception as e:  # noqa: BLE001
        err = str(e)
    ok = torch.tensor([1 if comm is not None else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if int(ok.item()) == 1:
        return comm, ("fused into the decrypt kernel: its last CTA stores the count into every rank's mailbox over "
                      "NVLink and the launch closing the batch collects the sums (csrc/peer.cuh); no NCCL call in the step")
    sys.stderr.write("bench: peer mailboxes unavailable (%s); using the NCCL all-reduce\n" % err)
    return None, "separate NCCL all-reduce (peer mailboxes unavailable: %s)" % (err or "on another rank")

def run_ours(args):
    import torch
    import torch.distributed as dist
    from csgn_b200 import engine as eng

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    N, D, T1, T2, desc = WORKLOADS[args.workload]
    T1g, T2 = args.t1 or T1, args.t2 or T2          # T1g: the left operand as given on the command line
    if args.t1 or args.t2:
        desc = "Context(%d,%d): %dx%d -> %d output blocks, multiply then decrypt" % (N, D, T1g, T2, T1g * T2)
    L, P = words_per_block(N), args.pairs
    eng.init(local)
    # a real (non-default) stream: handle 0 would mean "the library's own stream" to csgn_set_stream,
    # and torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    ctx = eng.Context(N, D)
    comm, exchange = connect_exchange(args, world, dev)

    # --- synthetic inputs: pinned host copies (e2e) and device copies (value) ---------
    # Global left operand of pair p has T1*world blocks; this rank owns csgn_shard_range.
    strong = args.scaling == "strong" and world > 1
    first, count = eng.shard_range(T1g if strong else T1g * world, rank, world)
    T1 = count                                        # this rank's blocks of the left operand
    if strong and T1g % world:
        raise SystemExit("--scaling strong needs --t1 divisible by the number of GPUs")
    host_a = torch.empty((P, T1 * L), dtype=torch.int64).pin_memory()
    host_b = torch.empty((P, T2 * L), dtype=torch.int64).pin_memory()
    key_pos = np.random.default_rng(7).permutation(N)[:D].astype(np.uint64)
    key = eng.SecretKey(ctx, key_pos)
    key_mask = np.zeros(L, dtype=np.uint64)
    for s_ in key_pos:
        key_mask[int(s_) >> 6] |= np.uint64(1 << (63 - (int(s_) & 63)))

    def planted(rng, T):
        # raw random blocks almost never satisfy a D=16 key; set the key bits in a few
        # of them so that the decrypt fold has something to count
        w = seeded_blocks(rng, T, N).reshape(T, L)
        rows = rng.choice(T, size=min(T, int(rng.integers(20, 60))), replace=False)
        w[rows] |= key_mask
        return w.reshape(-1)

    for p in range(P):
        rng_a = np.random.default_rng([1000 + p, rank])       # this rank's shard of A_p
        rng_b = np.random.default_rng([2000 + p])             # B_p, identical on every rank
        host_a[p].numpy().view(np.uint64)[:] = planted(rng_a, T1)
        host_b[p].numpy().view(np.uint64)[:] = planted(rng_b, T2)
    dev_a, dev_b = host_a.to(dev), host_b.to(dev)
    out = torch.empty((P, T1 * T2 * L), dtype=torch.int64, device=dev)
    # two result buffers: the all-reduce of step k runs on NCCL's stream while step k+1 computes
    counts2 = [torch.zeros(P, dtype=torch.int64, device=dev) for _ in range(2)]
    counts = counts2[0]
    pending = [None, None]
    host_counts = torch.zeros(P, dtype=torch.int64).pin_memory()
    va = [eng.Ciphertext.from_tensor(dev_a[p], ctx) for p in range(P)]
    vb = [eng.Ciphertext.from_tensor(dev_b[p], ctx) for p in range(P)]
    vo = [eng.Ciphertext.from_tensor(out[p], ctx) for p in range(P)]
    count_ptrs2 = [[c.data_ptr() + 8 * p for p in range(P)] for c in counts2]
    count_ptrs = count_ptrs2[0]
    step_no = [0]

    def drain():
        for i in (0, 1):
            if pending[i] is not None:
                pending[i].wait()
                pending[i] = None

    lagged = comm is not None and args.collect == "lagged"
    use_batch = args.enqueue == "batch"
    arr_a, arr_b, arr_o = eng.handle_array(va), eng.handle_array(vb), eng.handle_array(vo)
    S = max(1, min(args.streams, P))
    streams = [stream] + [torch.cuda.Stream(device=dev) for _ in range(S - 1)]
    sptr = [s_.cuda_stream for s_ in streams]
    fork_ev = [torch.cuda.Event() for _ in range(4)]
    join_ev = [[torch.cuda.Event() for _ in range(S)] for _ in range(4)]

    def fork(i):
        """side streams wait for what the main stream holds so far"""
        if S > 1:
            fork_ev[i].record(stream)
            for s_ in streams[1:]:
                s_.wait_event(fork_ev[i])

    def join(i):
        """the main stream waits for the side streams"""
        for k in range(1, S):
            join_ev[i][k].record(streams[k])
            stream.wait_event(join_ev[i][k])
        if S > 1:
            eng.set_stream(sptr[0])

    def step_device(evs=None, final=True):
        slot = step_no[0] & 1
        step_no[0] += 1
        if pending[slot] is not None:        # the all-reduce issued two steps ago: long finished
            pending[slot].wait()
            pending[slot] = None
        if evs:
            evs[0].record()
        if use_batch:
            eng.mul_into_batch(None, None, None, arrays=(arr_a, arr_b, arr_o))
            if evs:
                evs[1].record()
            if comm is None:
                key.count_satisfied_batch_async(None, count_ptrs2[slot][0], array=arr_o)
            elif lagged and step_no[0] > 1:
                comm.push_batch(key, None, count_ptrs2[slot ^ 1][0], lag=P, array=arr_o)
                if final:
                    comm.collect(P, count_ptrs2[slot][0])
            else:
                comm.push_batch(key, None, count_ptrs2[slot][0], array=arr_o)
            if evs:
                evs[2].record()
            if world > 1 and comm is None:
                pending[slot] = dist.all_reduce(counts2[slot], async_op=True)
            if evs:
                evs[3].record()
            return
        fork(0)
        for p in range(P):
            if S > 1:
                eng.set_stream(sptr[p % S])
            va[p].mul_into(vb[p], vo[p])
        join(0)
        if evs:
            evs[1].record()
        fork(1)
        if comm is not None:
            for p in range(P - 1):
                if S > 1:
                    eng.set_stream(sptr[p % S])
                comm.push(key, vo[p])                # fold; the count stays in the rank's local ring
            join(1)
            # the launch that closes the batch publishes the P counts to every rank and collects P sums:
            # this step's, or (lagged) the previous step's, which have long arrived
            if lagged and step_no[0] > 1:
                comm.push(key, vo[P - 1], P, count_ptrs2[slot ^ 1][0], lag=P)
                if final:                            # nothing follows: fetch this step's sums too
                    comm.collect(P, count_ptrs2[slot][0])
            else:
                comm.push(key, vo[P - 1], P, count_ptrs2[slot][0])
        else:
            for p in range(P):
                if S > 1:
                    eng.set_stream(sptr[p % S])
                key.count_satisfied_async(vo[p], count_ptrs2[slot][p])
            join(1)
        if evs:
            evs[2].record()
        if world > 1 and comm is None:
            pending[slot] = dist.all_reduce(counts2[slot], async_op=True)
        if evs:
            evs[3].record()

    def barrier():
        drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W_ = max(3, args.warmup)
    for i in range(W_):
        step_device(final=(i == W_ - 1))
    barrier()

    # sanity (outside the timed region, no oracle): satisfied-block counts are multiplicative
    ca = [key.count_satisfied(va[p]) for p in range(P)]
    cb = [key.count_satisfied(vb[p]) for p in range(P)]
    local_counts = torch.tensor([a * b for a, b in zip(ca, cb)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(local_counts)
    for c in counts2:
        if not torch.equal(local_counts, c):
            raise SystemExit("bench sanity failed: count(a*b) != count(a)*count(b): %s vs %s" % (c, local_counts))

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler = ClockSampler(local)
    sampler.start()                      # before the barrier: NVML's first calls are slow
    barrier()
    launches0 = eng.launch_count()
    t_begin = time.perf_counter()
    for k in range(K):
        step_device(evs[k], final=(k == K - 1))
    barrier()
    t_end = time.perf_counter()
    t_wall = t_end - t_begin
    clocks = sampler.stop(t_begin, t_end)
    launches = eng.launch_count() - launches0

    total_ms = evs[0][0].elapsed_time(evs[K - 1][3])
    mul_steps = [e[0].elapsed_time(e[1]) for e in evs]
    dec_steps = [e[1].elapsed_time(e[2]) for e in evs]
    mul_ms, dec_ms = sum(mul_steps), sum(dec_steps)
    ar_ms = sum(e[2].elapsed_time(e[3]) for e in evs)
    t = torch.tensor([total_ms, mul_ms, dec_ms, ar_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, mul_ms, dec_ms, ar_ms = (float(x) for x in t.tolist())

    blocks_per_step = P * T1 * T2 * world            # whole job (T1 = per-rank share)
    value = blocks_per_step * K / (total_ms * 1e-3)
    bytes_per_block = 8 * L
    peak, peak_src = measured_peak_gbs()
    mul_gbs = P * T1 * T2 * bytes_per_block * K / (mul_ms * 1e-3) / 1e9      # per GPU
    dec_gbs = P * T1 * T2 * bytes_per_block * K / (dec_ms * 1e-3) / 1e9

    # --- e2e: pinned host operands through the public C ABI ---------------------------
    e2e = None
    if not args.no_e2e:
        a_ptrs = [host_a[p].data_ptr() for p in range(P)]
        b_ptrs = [host_b[p].data_ptr() for p in range(P)]

        want_host = local_counts.cpu()
        host_counts2 = [torch.zeros(P, dtype=torch.int64).pin_memory() for _ in range(2)]
        done_ev = [torch.cuda.Event(), torch.cuda.Event()]
        in_flight = [False, False]
        e2e_no = [0]

        def check(slot):
            """the host reads step `slot`'s result (waits for its D2H) and compares it"""
            if in_flight[slot]:
                done_ev[slot].synchronize()
                in_flight[slot] = False
                if not torch.equal(host_counts2[slot], want_host):
                    raise SystemExit("e2e result differs from the device-resident result: %s" % host_counts2[slot])

        def step_e2e():
            # Depth-2 pipeline: step k is enqueued in full (uploads, kernels, D2H of its counts) BEFORE the host waits
            # for step k-1's result, so the GPU never idles while the host reads and checks.  Every step's result is
            # still read and verified on the host; the timed region ends after the last one has been.
            slot = e2e_no[0] & 1
            e2e_no[0] += 1
            check(slot)                                                    # frees this slot's buffers (step k-2)
            if use_batch and args.e2e_enqueue == "batch":
                has = [eng.Ciphertext.from_host_ptr(a_ptrs[p], T1, ctx) for p in range(P)]   # H2D on the copy stream
                hbs = [eng.Ciphertext.from_host_ptr(b_ptrs[p], T2, ctx) for p in range(P)]
                prods = eng.mul_batch(has, hbs)                            # csgn_mul_batch (allocates the P products)
                if comm is not None:
                    comm.push_batch(key, prods, count_ptrs2[slot][0])
                else:
                    key.count_satisfied_batch_async(prods, count_ptrs2[slot][0])
                del has, hbs, prods
                if world > 1 and comm is None:
                    dist.all_reduce(counts2[slot])
                host_counts2[slot].copy_(counts2[slot], non_blocking=True)
                done_ev[slot].record(stream)
                in_flight[slot] = True
                check(slot ^ 1)
                return
            fork(2)
            for p in range(P):
                if S > 1:
                    if comm is not None and p == P - 1:
                        join(2)                                            # the closing launch follows every push
                    else:
                        eng.set_stream(sptr[p % S])
                ha = eng.Ciphertext.from_host_ptr(a_ptrs[p], T1, ctx)      # H2D, async (pinned), on the copy stream
                hb = eng.Ciphertext.from_host_ptr(b_ptrs[p], T2, ctx)
                prod = ha * hb                                             # csgn_mul (allocates)
                if comm is not None:
                    comm.push(key, prod, P if p == P - 1 else 0, count_ptrs2[slot][0])
                else:
                    key.count_satisfied_async(prod, count_ptrs2[slot][p])
                del ha, hb, prod                                           # stream-ordered frees (on the pair's stream)
            if comm is None:
                join(2)
            if world > 1 and comm is None:
                dist.all_reduce(counts2[slot])
            host_counts2[slot].copy_(counts2[slot], non_blocking=True)     # D2H of this step's result
            done_ev[slot].record(stream)
            in_flight[slot] = True
            check(slot ^ 1)                                                # the previous step's result, now

        def drain_e2e():
            check(0)
            check(1)

        for _ in range(max(3, args.warmup)):
            step_e2e()
        drain_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_e2e()
        drain_e2e()                                                        # the last result is on the host ...
        e1.record()                                                        # ... before the clock stops
        barrier()
        e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
        e2e = {"value": blocks_per_step * K / (float(e2e_ms.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int(P * (T1 + T2) * L * 8), "d2h_bytes_per_step": int(P * 8),
               "ms_per_step": float(e2e_ms.item()) / K,
               "path": (("csgn_buf_upload x2P (pinned host) -> csgn_mul_batch -> %s"
                         if use_batch and args.e2e_enqueue == "batch" else
                         "csgn_buf_upload x2 (pinned host) -> csgn_mul -> %s")
                        % (("csgn_decrypt_sharded" if comm is not None else "csgn_decrypt_count") +
                           ("_batch_async" if use_batch and args.e2e_enqueue == "batch" else "_async")))
                       + "; one D2H of the P counts per step, read and checked on the host every step (one step "
                         "behind the enqueue); per GPU"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(3, args.warmup),
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": args.workload + ": " + desc, "pairs_per_step": P, "enqueue": args.enqueue,
                       "streams": ("library lanes (CSGN_LANES, default 2)" if use_batch else S),
                       "e2e_enqueue": (args.e2e_enqueue if use_batch else "pairs") + (" on %d streams" % S if not (use_batch and args.e2e_enqueue == "batch") else ""),
                       "left_blocks_per_rank": T1, "right_blocks": T2,
                       "blocks_per_step": blocks_per_step, "bytes_per_block": bytes_per_block,
                       "l2": "no flush needed: a step writes then reads %d x %.0f MB of products (>> 126 MB L2), "
                             "each product is read %d kernels after it was written" % (P, T1 * T2 * L * 8 / 1e6, P),
                       "sharding": "left operand by block range, right operand replicated; per step one %d-word "
                                   "exchange of the counts" % P if world > 1 else "single GPU",
                       "exchange": exchange + ((" [collect: %s]" % args.collect) if comm is not None else ""),
                       "inputs": "numpy default_rng raw blocks, pad bits zero, key bits set in 20-60 blocks per operand; "
                                 "key = default_rng(7).permutation(N)[:D]"},
            "roofline": {"bound": "hbm", "kernel": "mul_outer_kernel", "achieved": mul_gbs, "peak": peak,
                         "unit": "GB/s", "frac": mul_gbs / peak, "frac_of_nominal_8000": mul_gbs / 8000.0,
                         "traffic": ncu_traffic("mul_outer", args.workload),
                         "traffic_note": "ncu --set full, one isolated cold-cache launch: dram read+write bytes; the rest "
                                         "of the 160 MB product is still dirty in the 126 MB L2 when the launch ends and "
                                         "is written back under the next kernel (profiles/README.md)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": T1 * T2 * bytes_per_block,
                         "avg_launch_us": mul_ms * 1e3 / (K * P)},
            "kernels": {"multiply": {"blocks_per_s_per_gpu": P * T1 * T2 * K / (mul_ms * 1e-3), "gbs": mul_gbs,
                                     "frac_of_peak": mul_gbs / peak, "avg_launch_us": mul_ms * 1e3 / (K * P),
                                     "min_step_launch_us": min(mul_steps) * 1e3 / P,
                                     "median_step_launch_us": float(np.median(mul_steps)) * 1e3 / P},
                        "decrypt": {"blocks_per_s_per_gpu": P * T1 * T2 * K / (dec_ms * 1e-3), "gbs": dec_gbs,
                                    "frac_of_peak": dec_gbs / peak, "avg_launch_us": dec_ms * 1e3 / (K * P),
                                    "min_step_launch_us": min(dec_steps) * 1e3 / P,
                                    "median_step_launch_us": float(np.median(dec_steps)) * 1e3 / P},
                        "allreduce_ms_per_step": ar_ms / K},
            "clocks": clocks, "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * t_wall / K,
        }
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_extras:
            line["other_kernels"] = other_kernels(eng, torch, ctx, vo, N, D, L, peak, not args.no_cpu_baseline)
        if world == 1 and not args.no_cpu_baseline:
            # the reference overflows its int counters above 131,080 blocks at N=16383 (SURVEY hazard 4): the CPU
            # sample always uses the workload's own sizes, whatever --t1/--t2 say
            line["cpu_baseline"] = cpu_baseline(N, D, *WORKLOADS[args.workload][2:4])
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def run_chain(args):
    """cfg4: x = a_shard * b (1M blocks), y = x * d (1e6*chain_d blocks), decrypt(y), all-reduce."""
    import torch
    import torch.distributed as dist
    from csgn_b200 import engine as eng

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N, D, T1, T2, desc = WORKLOADS["cfg4"]
    Td, L = args.chain_d, words_per_block(N)
    eng.init(local)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    ctx = eng.Context(N, D)
    comm, exchange = connect_exchange(args, world, dev)
    key_pos = np.random.default_rng(7).permutation(N)[:D].astype(np.uint64)
    key = eng.SecretKey(ctx, key_pos)
    key_mask = np.zeros(L, dtype=np.uint64)
    for s_ in key_pos:
        key_mask[int(s_) >> 6] |= np.uint64(1 << (63 - (int(s_) & 63)))

    def planted(rng, T, k):
        w = seeded_blocks(rng, T, N).reshape(T, L)
        w[rng.choice(T, size=min(T, k), replace=False)] |= key_mask
        return w.reshape(-1)

    first, count = eng.shard_range(T1 * world, rank, world)
    host = {"a": planted(np.random.default_rng([1, rank]), T1, 31), "b": planted(np.random.default_rng([2]), T2, 17),
            "d": planted(np.random.default_rng([3]), Td, 5)}
    pinned = {k: torch.from_numpy(v.view(np.int64)).pin_memory() for k, v in host.items()}
    devt = {k: v.to(dev) for k, v in pinned.items()}
    va, vb, vd = (eng.Ciphertext.from_tensor(devt[k], ctx) for k in ("a", "b", "d"))
    x = torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev)
    y = torch.empty(T1 * T2 * Td * L, dtype=torch.int64, device=dev)
    vx, vy = eng.Ciphertext.from_tensor(x, ctx), eng.Ciphertext.from_tensor(y, ctx)
    counts = torch.zeros(1, dtype=torch.int64, device=dev)
    host_counts = torch.zeros(1, dtype=torch.int64).pin_memory()

    def step(evs=None):
        if evs:
            evs[0].record()
        va.mul_into(vb, vx)
        if evs:
            evs[1].record()
        vx.mul_into(vd, vy)
        if evs:
            evs[2].record()
        if comm is not None:
            comm.push(key, vy, 1, counts.data_ptr())     # fold, push and collect in the one kernel
        else:
            key.count_satisfied_async(vy, counts.data_ptr())
        if evs:
            evs[3].record()
        if world > 1 and comm is None:
            dist.all_reduce(counts)
        if evs:
            evs[4].record()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    W = max(3, args.warmup)
    for _ in range(W):
        step()
    barrier()
    want = torch.tensor([key.count_satisfied(va) * key.count_satisfied(vb) * key.count_satisfied(vd)],
                        dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(want)
    if not torch.equal(want, counts):
        raise SystemExit("cfg4 sanity failed: count((a*b)*d) %s != count(a)count(b)count(d) %s" % (counts, want))

    K = args.steps
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(K)]
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    launches0 = eng.launch_count()
    t_begin = time.perf_counter()
    for k in range(K):
        step(evs[k])
    barrier()
    clocks = sampler.stop(t_begin, time.perf_counter())
    launches = eng.launch_count() - launches0
    phases = [sum(e[i].elapsed_time(e[i + 1]) for e in evs) for i in range(4)]
    t = torch.tensor([evs[0][0].elapsed_time(evs[K - 1][4])] + phases, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, mul1_ms, mul2_ms, dec_ms, ar_ms = (float(v) for v in t.tolist())
    out_blocks = T1 * T2 * Td                      # per GPU
    value = out_blocks * world * K / (total_ms * 1e-3)
    peak, peak_src = measured_peak_gbs()
    mul2_gbs = out_blocks * 8 * L * K / (mul2_ms * 1e-3) / 1e9
    dec_gbs = out_blocks * 8 * L * K / (dec_ms * 1e-3) / 1e9

    # e2e: the three operands come from pinned host memory every step; the bit goes back
    def step_e2e():
        ha = eng.Ciphertext.from_host_ptr(pinned["a"].data_ptr(), T1, ctx)
        hb = eng.Ciphertext.from_host_ptr(pinned["b"].data_ptr(), T2, ctx)
        hd = eng.Ciphertext.from_host_ptr(pinned["d"].data_ptr(), Td, ctx)
        ha.mul_into(hb, vx)
        vx.mul_into(hd, vy)          # the 20 GB product is written in place: no room for two of them
        if comm is not None:
            comm.push(key, vy, 1, counts.data_ptr())
        else:
            key.count_satisfied_async(vy, counts.data_ptr())
        if world > 1 and comm is None:
            dist.all_reduce(counts)
        host_counts.copy_(counts, non_blocking=True)
        stream.synchronize()

    e2e = None
    if not args.no_e2e:
        for _ in range(W):
            step_e2e()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(K):
            step_e2e()
        e1.record()
        barrier()
        ems = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ems, op=dist.ReduceOp.MAX)
        if int(host_counts.item()) != int(want.item()):
            raise SystemExit("cfg4 e2e result differs")
        e2e = {"value": out_blocks * world * K / (float(ems.item()) * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": int((T1 + T2 + Td) * L * 8), "d2h_bytes_per_step": 8,
               "ms_per_step": float(ems.item()) / K,
               "path": "csgn_buf_upload x3 (pinned host) -> csgn_mul_into x2 -> csgn_decrypt_count_async -> D2H count"}
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": "cfg4: " + desc, "chain_d": Td, "blocks_per_gpu": out_blocks,
                           "blocks_total": out_blocks * world, "product_bytes_per_gpu": out_blocks * 8 * L,
                           "l2": "no flush needed: the product (%.1f GB per GPU) is far larger than L2" % (out_blocks * 8 * L / 1e9),
                           "sharding": "left operand by block range, b and d replicated, one-word exchange per decrypt",
                           "exchange": exchange},
                "roofline": {"bound": "hbm", "kernel": "mul_outer_kernel", "achieved": mul2_gbs, "peak": peak, "unit": "GB/s",
                             "frac": mul2_gbs / peak, "traffic": None, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": out_blocks * 8 * L, "avg_launch_us": mul2_ms * 1e3 / K},
                "kernels": {"multiply_1M": {"avg_launch_us": mul1_ms * 1e3 / K},
                            "multiply_chain": {"gbs": mul2_gbs, "frac_of_peak": mul2_gbs / peak, "avg_launch_us": mul2_ms * 1e3 / K},
                            "decrypt": {"gbs": dec_gbs, "frac_of_peak": dec_gbs / peak, "avg_launch_us": dec_ms * 1e3 / K},
                            "allreduce_ms_per_step": ar_ms / K},
                "clocks": clocks, "gpu_launches": int(launches)}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_baseline(N, D, T1, T2)
            cb["sample"] += "; per-block cost of the chain is the same multiply+decrypt work"
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload == "cfg4":
        return run_chain(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
