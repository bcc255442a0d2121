"""Time the permute kernels (bit-sliced vs word-gather) on the GPU box."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
for N, T, nbuf in ((1247, 1000000, 6), (16383, 90000, 6), (1247, 10000000, 2)):
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
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for i in range(nbuf): vi[i].permute_into(perm, vo[i])
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * nbuf)
    for label, env in (("fixed (default)", {}), ("fixed waves=2", {"CSGN_PERM_WAVES": "2"}), ("fixed waves=4", {"CSGN_PERM_WAVES": "4"}),
                       ("fixed waves=16", {"CSGN_PERM_WAVES": "16"}),
                       ("fixed variant 2", {"CSGN_PERM_VARIANT": "2"}), ("fixed variant 2 waves=4", {"CSGN_PERM_VARIANT": "2", "CSGN_PERM_WAVES": "4"}),
                       ("fixed variant 3", {"CSGN_PERM_VARIANT": "3"}),
                       ("prefetch 2 buffers", {"CSGN_PERM_VARIANT": "4"}), ("prefetch 1 buffer", {"CSGN_PERM_VARIANT": "5"}),
                       ("prefetch 2 buffers waves=4", {"CSGN_PERM_VARIANT": "6"}), ("prefetch 1 buffer waves=4", {"CSGN_PERM_VARIANT": "7"}),
                       ("prefetch 1 buffer waves=2", {"CSGN_PERM_VARIANT": "5", "CSGN_PERM_WAVES": "2"}),
                       ("register-fed (variant 8)", {"CSGN_PERM_VARIANT": "8"}),
                       ("runtime-W kernel", {"CSGN_PERM_VARIANT": "1"}), ("runtime-W waves=1", {"CSGN_PERM_VARIANT": "1", "CSGN_PERM_WAVES": "1"}),
                       ("gather", {"CSGN_PERM_GATHER": "1"})):
        for k in ("CSGN_PERM_ITEMS", "CSGN_PERM_WAVES", "CSGN_PERM_GATHER", "CSGN_PERM_VARIANT"): os.environ.pop(k, None)
        os.environ.update(env)
        ms = timed()
        print("N=%d T=%d %-28s %9.2f us  %7.1f GB/s (read+write)  %.3g blocks/s" % (N, T, label, ms * 1e3, nbytes / ms / 1e6, T / ms * 1e3), flush=True)
    del ins, outs, vi, vo
    torch.cuda.empty_cache()
