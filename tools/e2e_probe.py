import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from csgn_b200 import engine as eng
S = int(sys.argv[1]); depth = int(sys.argv[2])
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
N, D, T = 1247, 16, 1000; ctx = eng.Context(N, D); L = ctx.L; P = 16
host_a = torch.randint(-2**62, 2**62, (P, T * L), dtype=torch.int64).pin_memory()
host_b = torch.randint(-2**62, 2**62, (P, T * L), dtype=torch.int64).pin_memory()
key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D].astype(np.uint64))
counts = [torch.zeros(P, dtype=torch.int64, device=dev) for _ in range(2)]
hc = [torch.zeros(P, dtype=torch.int64).pin_memory() for _ in range(2)]
ev = [torch.cuda.Event(), torch.cuda.Event()]; fl = [False, False]
streams = [stream] + [torch.cuda.Stream() for _ in range(S - 1)]
fe = torch.cuda.Event(); je = [torch.cuda.Event() for _ in range(S)]
def step(k):
    slot = k & 1
    if S > 1:
        fe.record(stream)
        for s_ in streams[1:]: s_.wait_event(fe)
    for p in range(P):
        if S > 1: eng.set_stream(streams[p % S].cuda_stream)
        ha = eng.Ciphertext.from_host_ptr(host_a[p].data_ptr(), T, ctx)
        hb = eng.Ciphertext.from_host_ptr(host_b[p].data_ptr(), T, ctx)
        prod = ha * hb
        key.count_satisfied_async(prod, counts[slot].data_ptr() + 8 * p)
        del ha, hb, prod
    for i in range(1, S):
        je[i].record(streams[i]); stream.wait_event(je[i])
    eng.set_stream(stream.cuda_stream)
    hc[slot].copy_(counts[slot], non_blocking=True)
    ev[slot].record(stream); fl[slot] = True
    other = slot ^ 1 if depth == 2 else slot
    if fl[other]: ev[other].synchronize(); fl[other] = False
times = []
for k in range(40):
    t0 = time.perf_counter(); step(k); times.append((time.perf_counter() - t0) * 1e3)
torch.cuda.synchronize()
info = eng.device_info()
print("S=%d depth=%d host ms/step: first10 %s  last10 %s  reserved %.2f GB" % (S, depth, [round(t, 2) for t in times[:10]], [round(t, 2) for t in times[-10:]], (info["hbm_total"] - info["hbm_free"]) / 1e9))
