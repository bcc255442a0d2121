import os, sys
sys.path.insert(0, os.getcwd())
os.environ["CSGN_TUNING"]="1"
import numpy as np, torch
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
for N, T, nbuf in ((4097, 313600, 6), (6200, 200000, 6), (12351, 104000, 6), (10000, 128000, 6), (5000, 256000, 6)):
    ctx = eng.Context(N, 16); L = ctx.L
    ins = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(nbuf)]
    outs = [torch.empty(T * L, dtype=torch.int64, device=dev) for _ in range(nbuf)]
    vi = [eng.Ciphertext.from_tensor(t, ctx) for t in ins]; vo = [eng.Ciphertext.from_tensor(t, ctx) for t in outs]
    perm = eng.Permutation(ctx, np.random.default_rng(3).permutation(N))
    for label, env in (("default", {}), ("128x6", {"CSGN_PERM_PLANE": "17"}), ("192x4", {"CSGN_PERM_PLANE": "18"}), ("256x3", {"CSGN_PERM_PLANE": "7"})):
        os.environ.pop("CSGN_PERM_PLANE", None); os.environ.update(env)
        for i in range(nbuf): vi[i].permute_into(perm, vo[i])
        torch.cuda.synchronize()
        res = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                for i in range(nbuf): vi[i].permute_into(perm, vo[i])
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / (3 * nbuf))
        ms = float(np.median(res))
        print("N=%-6d W=%-4d %-8s %9.2f us  %.3f of peak" % (N, 2 * L, label, ms * 1e3, 2 * T * L * 8 / ms / 1e6 / 6533.2), flush=True)
    del ins, outs, vi, vo; torch.cuda.empty_cache()
