"""Diagnose per-launch time of the bench pattern under a few conditions (GPU box)."""
import os, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csgn_b200 import engine as eng

N, D, T1, T2, P = 1247, 16, 1000, 1000, 16
torch.cuda.set_device(0); dev = torch.device("cuda", 0)
eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
ctx = eng.Context(N, D); L = ctx.L
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randint(-2**62, 2**62, (P, T1 * L), dtype=torch.int64, device=dev, generator=g)
B = torch.randint(-2**62, 2**62, (P, T2 * L), dtype=torch.int64, device=dev, generator=g)
key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D])
cnt = torch.zeros(P, dtype=torch.int64, device=dev)

def run(label, outs, distinct_ops, sampler=False, interleave=False, K=30):
    va = [eng.Ciphertext.from_tensor(A[p if distinct_ops else 0], ctx) for p in range(P)]
    vb = [eng.Ciphertext.from_tensor(B[p if distinct_ops else 0], ctx) for p in range(P)]
    vo = [eng.Ciphertext.from_tensor(o, ctx) for o in outs]
    stop = threading.Event()
    def poll():
        import pynvml
        pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
        while not stop.is_set():
            pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksEventReasons(h); time.sleep(0.002)
    th = threading.Thread(target=poll, daemon=True)
    def step(ev=None):
        if ev: ev[0].record()
        if interleave:
            for p in range(P):
                va[p].mul_into(vb[p], vo[p]); key.count_satisfied_async(vo[(p + P // 2) % P], cnt.data_ptr() + 8 * p)
            if ev: ev[1].record(); ev[2].record()
        else:
            for p in range(P): va[p].mul_into(vb[p], vo[p])
            if ev: ev[1].record()
            for p in range(P): key.count_satisfied_async(vo[p], cnt.data_ptr() + 8 * p)
            if ev: ev[2].record()
    for _ in range(5): step()
    torch.cuda.synchronize()
    if sampler: th.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    t0 = time.perf_counter()
    for k in range(K): step(evs[k])
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    stop.set()
    mul = np.median([e[0].elapsed_time(e[1]) for e in evs]) * 1e3 / P
    dec = np.median([e[1].elapsed_time(e[2]) for e in evs]) * 1e3 / P
    tot = evs[0][0].elapsed_time(evs[-1][2]) * 1e3 / (K * P)
    print("%-46s mul %6.2f us  dec %6.2f us  pair %6.2f us  (cpu enqueue %5.1f us/launch)" % (label, mul, dec, tot, t_enq * 1e6 / (K * P * 2)), flush=True)

sep = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(P)]
for var in (0, 1, 2, 5, 6, 7):
    os.environ["CSGN_DEC_VARIANT"] = str(var)
    run("decrypt variant %d" % var, sep, True)
