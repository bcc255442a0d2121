"""Window walk (odd-L decrypt): forms <loads in flight, CTAs per SM, refills per group> and residency caps at 160 MB (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CSGN_TUNING"] = "1"
import numpy as np, torch
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
PEAK = 6533.2
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
for N, T in ((191, 9000000), (1215, 1000000), (4097, 313600), (12351, 102400), (4097, 3136000)):
    ctx = eng.Context(N, 16); L = ctx.L; P = 8 if T * L * 8 < (1 << 30) else 2
    A = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(P)]
    va = [eng.Ciphertext.from_tensor(x, ctx) for x in A]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:16].astype(np.uint64))
    cnt = torch.zeros(P, dtype=torch.int64, device=dev)
    nb = T * L * 8
    names = {3: "<8,3,1>", 2: "<6,4,1>", 4: "<12,2,1>", 6: "<8,3,2>", 11: "<6,4,2>"}
    row = "N=%-6d L=%-4d |" % (N, L)
    for form in (3, 6, 2, 11, 4):
        os.environ["CSGN_DEC_WINDOW"] = str(form)
        d = timed(lambda i: key.count_satisfied_async(va[i], cnt.data_ptr() + 8 * i), P)
        row += " %s %.3f |" % (names[form], nb / d / 1e3 / PEAK)
    print(row, flush=True)
    for form, caps in ((3, (1, 2)), (2, (1, 2, 3)), (4, (1,))):
        row = "     CTAs per SM capped: %s |" % names[form]
        for cap in caps:
            os.environ["CSGN_DEC_WINDOW"] = str(form); os.environ["CSGN_DEC_CTAS_PER_SM"] = str(cap)
            d = timed(lambda i: key.count_satisfied_async(va[i], cnt.data_ptr() + 8 * i), P)
            row += " %d: %.3f |" % (cap, nb / d / 1e3 / PEAK)
        os.environ.pop("CSGN_DEC_CTAS_PER_SM", None)
        print(row, flush=True)
    os.environ.pop("CSGN_DEC_WINDOW", None); os.environ.pop("CSGN_DEC_WAVES", None)
    del A, va; torch.cuda.empty_cache()
