import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csgn_b200 import engine as eng
eng.init(0)
for N, D, n in ((1247, 16, 1000000), (16383, 64, 100000)):
    ctx = eng.Context(N, D)
    key = eng.SecretKey(ctx, np.random.default_rng(1).permutation(N)[:D])
    bits = np.random.default_rng(2).integers(0, 2, size=n).astype(np.uint8)
    key.encrypt_batch(bits, 1)
    t0 = time.perf_counter()
    for r in range(5): ct = key.encrypt_batch(bits, r)
    eng.sync(); dt = (time.perf_counter() - t0) / 5
    print("N=%d: %d fresh blocks in %.3f ms = %.3g blocks/s (%.0f GB/s written, incl. H2D of the bits and sync)" % (N, n, dt * 1e3, n / dt, n * ctx.L * 8 / dt / 1e9))
