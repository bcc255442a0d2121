"""Odd-L contexts: multiply and decrypt with 16-byte units over double blocks vs the 8-byte-unit kernels (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CSGN_TUNING"] = "1"
import numpy as np, torch
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
def timed(fn, n, reps=5):
    for i in range(n): fn(i)
    torch.cuda.synchronize()
    res = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / (reps * n))
    return float(np.median(res))
for N, T in ((4097, 560), (2111, 780), (1215, 1000), (6200, 460), (12351, 320)):
    ctx = eng.Context(N, 16); L = ctx.L; P = 8
    A = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(P)]
    B = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(P)]
    O = [torch.empty(T * T * L, dtype=torch.int64, device=dev) for _ in range(P)]
    va = [eng.Ciphertext.from_tensor(x, ctx) for x in A]; vb = [eng.Ciphertext.from_tensor(x, ctx) for x in B]
    vo = [eng.Ciphertext.from_tensor(x, ctx) for x in O]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:16].astype(np.uint64))
    cnt = torch.zeros(P, dtype=torch.int64, device=dev)
    nb = T * T * L * 8
    for label, env in (("double blocks", {}), ("string pairs", {"CSGN_DEC_PAIRS_MIN": "100000"}), ("rows pairs", {"CSGN_DEC_PAIRS_MIN": "17"}),
                       ("8-byte units", {"CSGN_MUL_DOUBLE": "0", "CSGN_DEC_PAIRS_MIN": "100000", "CSGN_DEC_STRING_PAIRS": "0"})):
        for k in ("CSGN_MUL_DOUBLE", "CSGN_DEC_PAIRS_MIN", "CSGN_DEC_STRING_PAIRS"): os.environ.pop(k, None)
        os.environ.update(env)
        m = timed(lambda i: va[i].mul_into(vb[i], vo[i]), P)
        d = timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr() + 8 * i), P)
        print("N=%-6d L=%-4d %-14s multiply %8.2f us %.3f | decrypt %8.2f us %.3f" % (N, L, label, m, nb / m / 1e3 / 6533.2, d, nb / d / 1e3 / 6533.2), flush=True)
    del A, B, O, va, vb, vo; torch.cuda.empty_cache()
