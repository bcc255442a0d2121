import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng
torch.cuda.set_device(0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
for N, D, n in ((1247, 16, 1000000), (16383, 64, 100000), (191, 5, 1000000)):
    ctx = eng.Context(N, D)
    key = eng.SecretKey(ctx, np.random.default_rng(1).permutation(N)[:D])
    bits = np.random.default_rng(2).integers(0, 2, size=n).astype(np.uint8)
    for r in range(3): ct = key.encrypt_batch(bits, r)
    ts = []
    for r in range(7):
        t0 = time.perf_counter(); ct = key.encrypt_batch(bits, r); eng.sync(); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print("N=%d: %d fresh blocks per call, median %.3f ms (min %.3f, max %.3f) = %.3g blocks/s incl. H2D of the bits, allocation and sync"
          % (N, n, dt * 1e3, min(ts) * 1e3, max(ts) * 1e3, n / dt))
