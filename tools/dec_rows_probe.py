import os, sys
sys.path.insert(0, os.getcwd())
os.environ["CSGN_TUNING"] = "1"
import numpy as np, torch
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
for N, T in ((33000, 40000), (33000, 400000), (12000, 120000)):
    ctx = eng.Context(N, 16); L = ctx.L
    bufs = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(8)]
    cts = [eng.Ciphertext.from_tensor(b, ctx) for b in bufs]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:16].astype(np.uint64))
    cnt = torch.zeros(8, dtype=torch.int64, device=dev)
    for label, env in (("rows", {}), ("rows bpi1", {"CSGN_DEC_ROWS_BPI": "1"}), ("rows bpi2", {"CSGN_DEC_ROWS_BPI": "2"}),
                       ("rows bpi4", {"CSGN_DEC_ROWS_BPI": "4"}), ("string", {"CSGN_DEC_ROWS_MIN": "100000"})):
        os.environ.pop("CSGN_DEC_ROWS_MIN", None); os.environ.pop("CSGN_DEC_ROWS_BPI", None); os.environ.update(env)
        for i in range(8): key.count_satisfied_async(cts[i], cnt.data_ptr() + 8 * i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            for i in range(8): key.count_satisfied_async(cts[i], cnt.data_ptr() + 8 * i)
        e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / 40
        print("N=%d L=%d T=%d %-10s %8.2f us %7.0f GB/s %.3f" % (N, L, T, label, us, T * L * 8 / us / 1e3, T * L * 8 / us / 1e3 / 6533.2), flush=True)
