import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csgn_b200 import engine as eng
N, D, T1, T2, P = 1247, 16, 1000, 1000, 16
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
ctx = eng.Context(N, D); L = ctx.L
g = torch.Generator(device=dev); g.manual_seed(1)
A = torch.randint(-2**62, 2**62, (P, T1 * L), dtype=torch.int64, device=dev, generator=g)
B = torch.randint(-2**62, 2**62, (P, T2 * L), dtype=torch.int64, device=dev, generator=g)
outs = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(P)]
vo = [eng.Ciphertext.from_tensor(o, ctx) for o in outs]
def run(label, distinct, K=20):
    va = [eng.Ciphertext.from_tensor(A[p if distinct else 0], ctx) for p in range(P)]
    vb = [eng.Ciphertext.from_tensor(B[p if distinct else 0], ctx) for p in range(P)]
    for _ in range(3):
        for p in range(P): va[p].mul_into(vb[p], vo[p])
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    for e0, e1 in evs:
        e0.record()
        for p in range(P): va[p].mul_into(vb[p], vo[p])
        e1.record()
    torch.cuda.synchronize()
    print("%-40s %6.2f us" % (label, float(np.median([a.elapsed_time(b) for a, b in evs])) * 1e3 / P), flush=True)
tag = os.environ.get("CSGN_LIBRARY", "current")
os.environ["CSGN_MUL_KERNEL"] = "1"
for pf in ("8", "0"):
    os.environ["CSGN_MUL_PF_CTAS_PER_SM"] = pf
    run(tag + " v1 pf=%s distinct" % pf, True); run(tag + " v1 pf=%s same" % pf, False)
