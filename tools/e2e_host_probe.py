"""Where does the HOST time of one e2e step go?  (cfg2, 16 pairs: 32 uploads, one fused batch call, 48 frees, one D2H)

    python tools/e2e_host_probe.py [steps]

perf_counter around each phase of the enqueue, no synchronisation inside the loop (the GPU runs behind); the last
line is the GPU-side time per step of the same loop for comparison."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csgn_b200 import engine as eng  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 300
N, D, T, P = 1247, 16, 1000, 16
torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
eng.init(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
ctx = eng.Context(N, D)
L = ctx.L
host_a = torch.randint(-2**62, 2**62, (P, T * L), dtype=torch.int64).pin_memory()
host_b = torch.randint(-2**62, 2**62, (P, T * L), dtype=torch.int64).pin_memory()
key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D].astype(np.uint64))
counts = torch.zeros(P, dtype=torch.int64, device=dev)
hcounts = torch.zeros(P, dtype=torch.int64).pin_memory()
a_ptrs = [host_a[p].data_ptr() for p in range(P)]
b_ptrs = [host_b[p].data_ptr() for p in range(P)]
acc = {"upload": 0.0, "arrays": 0.0, "batch_call": 0.0, "wrap": 0.0, "free": 0.0, "d2h": 0.0}


def step(timeit):
    t0 = time.perf_counter()
    has = [eng.Ciphertext.from_host_ptr(a_ptrs[p], T, ctx) for p in range(P)]
    hbs = [eng.Ciphertext.from_host_ptr(b_ptrs[p], T, ctx) for p in range(P)]
    t1 = time.perf_counter()
    ha, hb = eng.handle_array(has), eng.handle_array(hbs)
    ho = (ctypes.c_void_p * P)()
    t2 = time.perf_counter()
    eng.mul_count_batch_async(key, None, None, counts.data_ptr(), arrays=(ha, hb, ho))
    t3 = time.perf_counter()
    prods = [eng.Ciphertext(ctypes.c_void_p(ho[i]), ctx) for i in range(P)]
    t4 = time.perf_counter()
    del has, hbs, prods
    t5 = time.perf_counter()
    hcounts.copy_(counts, non_blocking=True)
    t6 = time.perf_counter()
    if timeit:
        for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
            acc[k] += v


for _ in range(20):
    step(False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
w0 = time.perf_counter()
for i in range(steps):
    step(True)
    if i % 8 == 7:
        stream.synchronize()          # keep the GPU queue short so that host timings are not queue back-pressure
w1 = time.perf_counter()
e1.record()
torch.cuda.synchronize()
tot = sum(acc.values())
for k, v in acc.items():
    print("  %-12s %8.1f us per step (%4.1f %%)" % (k, v / steps * 1e6, 100 * v / tot))
print("  host total   %8.1f us per step; wall %8.1f us; GPU events %8.1f us per step" %
      (tot / steps * 1e6, (w1 - w0) / steps * 1e6, e0.elapsed_time(e1) * 1e3 / steps))
