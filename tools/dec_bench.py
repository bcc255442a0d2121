"""Time the decrypt fold alone at several sizes and kernel variants (CSGN_DEC_VARIANT), for comparison with
tools/membw.cu's plain read stream.  Usage: python tools/dec_bench.py [variants...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
variants = [int(v) for v in sys.argv[1:]] or [0]
g = torch.Generator(device=dev); g.manual_seed(1)
for N, D, T, nbuf in ((1247, 16, 1000000, 16), (1247, 16, 25000000, 2), (16383, 64, 90000, 16), (16383, 64, 2000000, 2)):
    ctx = eng.Context(N, D); L = ctx.L
    bufs = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(nbuf)]
    cts = [eng.Ciphertext.from_tensor(t, ctx) for t in bufs]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D].astype(np.uint64))
    out = torch.zeros(nbuf, dtype=torch.int64, device=dev)
    for v in variants:
        os.environ["CSGN_DEC_VARIANT"] = str(v)
        def run():
            for i in range(nbuf): key.count_satisfied_async(cts[i], out.data_ptr() + 8 * i)
        run(); torch.cuda.synchronize()
        reps = 5 if T * L * 8 < 1e9 else 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run()
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nbuf)
        print("N=%d T=%d (%.2f GB) variant %d: %9.2f us  %7.1f GB/s" % (N, T, T * L * 8 / 1e9, v, us, T * L * 8 / us / 1e3), flush=True)
    del bufs, cts
    torch.cuda.empty_cache()
