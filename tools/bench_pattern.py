"""[mul x P][decrypt x P] with cold operands (the bench pattern) for several knob settings."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
N, D, T1, T2, P = 1247, 16, 1000, 1000, 16
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
ctx = eng.Context(N, D); L = ctx.L
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randint(-2**62, 2**62, (P, T1 * L), dtype=torch.int64, device=dev, generator=g)
B = torch.randint(-2**62, 2**62, (P, T2 * L), dtype=torch.int64, device=dev, generator=g)
key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D])
cnt = torch.zeros(P, dtype=torch.int64, device=dev)
outs = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(P)]
va = [eng.Ciphertext.from_tensor(A[p], ctx) for p in range(P)]
vb = [eng.Ciphertext.from_tensor(B[p], ctx) for p in range(P)]
vo = [eng.Ciphertext.from_tensor(o, ctx) for o in outs]
def run(label, K=30):
    def step(ev=None):
        if ev: ev[0].record()
        for p in range(P): va[p].mul_into(vb[p], vo[p])
        if ev: ev[1].record()
        for p in range(P): key.count_satisfied_async(vo[p], cnt.data_ptr() + 8 * p)
        if ev: ev[2].record()
    for _ in range(5): step()
    torch.cuda.synchronize()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for k in range(K): step(evs[k])
    torch.cuda.synchronize()
    mul = np.median([e[0].elapsed_time(e[1]) for e in evs]) * 1e3 / P
    dec = np.median([e[1].elapsed_time(e[2]) for e in evs]) * 1e3 / P
    print("%-40s mul %6.2f us  dec %6.2f us" % (label, mul, dec), flush=True)
def setenv(**kv):
    for k in list(os.environ):
        if k.startswith("CSGN_MUL_") or k.startswith("CSGN_DEC_") or k == "CSGN_PDL": del os.environ[k]
    for k, v in kv.items(): os.environ[k] = str(v)
setenv(); run("default")
setenv(CSGN_PDL=0); run("no PDL")
for cpw in (1, 2, 3, 4, 6, 8):
    setenv(CSGN_DEC_CPW=cpw); run("decrypt short CTAs, %d chunks/warp" % cpw)
    setenv(CSGN_DEC_CPW=cpw, CSGN_DEC_VARIANT=1); run("  same, 64 regs (4 CTAs/SM)")
    setenv(CSGN_DEC_CPW=cpw, CSGN_DEC_VARIANT=2); run("  same, half-chunk unroll (5 CTAs/SM)")
for c in (2, 3): setenv(CSGN_DEC_CTAS_PER_SM=c); run("decrypt <= %d CTAs/SM" % c)
setenv(CSGN_MUL_PF_CTAS_PER_SM=0); run("no L2 warm-up")
for ips in (16, 64): setenv(CSGN_MUL_ITEMS_PER_SM=ips); run("items/SM=%d" % ips)
