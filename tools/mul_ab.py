"""Multiply at several shapes (default heuristics, then forced units per thread), rotating operands. GPU box."""
import os, sys, itertools
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng

torch.cuda.set_device(0); dev = torch.device("cuda", 0)
eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
SHAPES = {"cfg2": (1247, 1000, 1000, 16), "cfg5": (16383, 300, 300, 14), "big": (1247, 5000, 5000, 2),
          "cfg5big": (16383, 1000, 1000, 3), "tallx1": (1247, 1000000, 1, 12), "1xwide": (1247, 1, 1000000, 12),
          "tallx7": (1247, 150000, 7, 12), "100x10k": (1247, 100, 10000, 12), "10kx100": (1247, 10000, 100, 12), "chain125": (1247, 1000000, 125, 1), "chain25": (1247, 1000000, 25, 2),
          "1Mx400": (1247, 1000000, 400, 1), "100kx1000": (1247, 100000, 1000, 1)}
which = sys.argv[1:] or list(SHAPES)
for name in which:
    N, T1, T2, P = SHAPES[name]
    ctx = eng.Context(N, 16); L = ctx.L
    A = torch.randint(-2**62, 2**62, (P, T1 * L), dtype=torch.int64, device=dev, generator=g)
    B = torch.randint(-2**62, 2**62, (P, T2 * L), dtype=torch.int64, device=dev, generator=g)
    outs = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(P)]
    va = [eng.Ciphertext.from_tensor(A[p], ctx) for p in range(P)]
    vb = [eng.Ciphertext.from_tensor(B[p], ctx) for p in range(P)]
    vo = [eng.Ciphertext.from_tensor(o, ctx) for o in outs]
    nbytes = T1 * T2 * L * 8
    def timed(K=12):
        for p in range(P): va[p].mul_into(vb[p], vo[p])
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
        for e0, e1 in evs:
            e0.record()
            for p in range(P): va[p].mul_into(vb[p], vo[p])
            e1.record()
        torch.cuda.synchronize()
        return float(np.median([a.elapsed_time(b) for a, b in evs])) / P
    res = []
    for kern, u in ((0, 0), (0, 1), (0, 2), (0, 4)):
        if u: os.environ["CSGN_MUL_U"] = str(u)
        else: os.environ.pop("CSGN_MUL_U", None)
        ms = timed()
        res.append("v%d/U%d %8.2f us %6.0f GB/s" % (kern, u, ms * 1e3, nbytes / ms / 1e6))
    ms = timed.__call__() if False else None
    z = outs[0]
    def zt():
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for o in outs: o.zero_()
        torch.cuda.synchronize(); e0.record()
        for _ in range(5):
            for o in outs: o.zero_()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (5 * P)
    zms = zt()
    print("%-8s %9.1f MB | %s | zero_ %8.2f us %6.0f GB/s" % (name, nbytes / 1e6, " | ".join(res), zms * 1e3, nbytes / zms / 1e6), flush=True)
    del A, B, outs, va, vb, vo
    torch.cuda.empty_cache()
