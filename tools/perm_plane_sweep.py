"""Time the plane permute kernel's instantiations against the current default (GPU box; CUDA events, rotating buffers).

    python tools/perm_plane_sweep.py
"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
PEAK = 6533.2
FORMS = {0: "<512,256,3,bulk>", 1: "<512,256,3,regs>", 2: "<512,512,1,bulk>", 3: "<512,512,2,regs>", 4: "<512,128,6,regs>",
         5: "<512,128,6,bulk>", 6: "<rt,256,3,regs>", 7: "<rt,256,3,bulk>", 8: "<rt,128,4,regs>", 9: "<rt,512,1,bulk>",
         10: "<512,512,2,bulk>", 11: "<512,384,2,bulk>", 12: "<256,128,6,bulk>", 13: "<256,256,3,bulk>", 14: "<256,256,4,bulk>",
         15: "<256,128,6,regs>", 16: "<rt,1024,1,bulk>", 17: "<rt,128,6,bulk>",
         20: "<40,G4,160,6>", 21: "<40,G4,160,4>", 22: "<40,G8,320,3>", 23: "<40,G2,96,10>", 24: "<40,G4,96,8>", 25: "<40,G6,256,4>"}
ONLY = [int(x) for x in os.environ.get("PLANE_FORMS", "").split(",") if x]
SHAPES = ((1247, 1000000, 6), (1247, 10000000, 2), (16383, 90000, 6), (16383, 1000000, 2), (8191, 160000, 6), (33000, 40000, 6), (4097, 313600, 6), (2048, 640000, 6))
if os.environ.get("PLANE_SHAPES"):
    SHAPES = SHAPES[:int(os.environ["PLANE_SHAPES"])]
for N, T, nbuf in SHAPES:
    ctx = eng.Context(N, 16); L = ctx.L
    ins = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(nbuf)]
    outs = [torch.empty(T * L, dtype=torch.int64, device=dev) for _ in range(nbuf)]
    vi = [eng.Ciphertext.from_tensor(t, ctx) for t in ins]
    vo = [eng.Ciphertext.from_tensor(t, ctx) for t in outs]
    perm = eng.Permutation(ctx, np.random.default_rng(3).permutation(N))
    nbytes = 2 * T * L * 8
    def timed(reps=3):
        for i in range(nbuf): vi[i].permute_into(perm, vo[i])
        torch.cuda.synchronize()
        res = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                for i in range(nbuf): vi[i].permute_into(perm, vo[i])
            e1.record(); torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / (reps * nbuf))
        return float(np.median(res))
    cases = [("default", {})]
    for f in FORMS:
        if ONLY and f not in ONLY:
            continue
        if (FORMS[f].startswith("<512") and 2 * L != 512) or (FORMS[f].startswith("<256") and 2 * L != 256) or \
                (FORMS[f].startswith("<40") != (2 * L == 40)):
            continue
        for w in [int(x) for x in os.environ.get('PLANE_WAVES', '1,2,4').split(',')]:
            cases.append(("plane %s waves=%d" % (FORMS[f], w), {"CSGN_PERM_PLANE": str(f), "CSGN_PERM_WAVES": str(w)}))
    ref = None
    for label, env in cases:
        for k in ("CSGN_PERM_PLANE", "CSGN_PERM_WAVES", "CSGN_PERM_VARIANT"): os.environ.pop(k, None)
        os.environ.update(env)
        ms = timed()
        if ref is None:
            ref = ms
        print("N=%-6d T=%-8d %-34s %9.2f us  %7.1f GB/s  %.3f of peak  (%+.1f %% vs default)" %
              (N, T, label, ms * 1e3, nbytes / ms / 1e6, nbytes / ms / 1e6 / PEAK, 100 * (ref / ms - 1)), flush=True)
    del ins, outs, vi, vo
    torch.cuda.empty_cache()
