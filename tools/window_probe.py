"""Odd-L decrypt: the window walk (csrc/decrypt.cu, decrypt_count_window_kernel) against the kernels it replaces (GPU box).

    python tools/window_probe.py            # every shape, every form
    python tools/window_probe.py one        # N=4097 only, default form (for ncu)
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CSGN_TUNING"] = "1"
import json
import numpy as np, torch
from csgn_b200 import engine as eng
torch.cuda.set_device(0); dev = torch.device("cuda", 0); eng.init(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
g = torch.Generator(device=dev); g.manual_seed(1)
_peaks = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
PEAK = float(json.load(open(_peaks)).get("hbm_gbs", 6533.2)) if os.path.exists(_peaks) else 6533.2
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
one = len(sys.argv) > 1 and sys.argv[1] == "one"
shapes = ((4097, 313600), (575, 2200000)) if one else ((191, 9000000), (319, 4000000), (447, 2900000), (575, 2200000), (703, 1800000), (831, 1500000),
                                        (959, 1300000), (1215, 1000000), (2111, 610000), (4097, 313600), (6200, 211600),
                                        (12351, 102400), (4097, 3136000),
                                        (2173, 590000), (2301, 550000), (3197, 400000), (3965, 320000), (6397, 200000), (12797, 100000), (16253, 78000))
forms = (("window", {}), ("before", {"CSGN_DEC_WINDOW": "0"})) if one else (("window (default form)", {"CSGN_DEC_WINDOW_ALL": "1"}), ("before", {"CSGN_DEC_WINDOW": "0"}))
print("# decrypt of odd-L ciphertexts, fraction of the measured copy peak (%.0f GB/s); 8 buffers in rotation" % PEAK)
for N, T in shapes:
    ctx = eng.Context(N, 16); L = ctx.L
    P = 8 if T * L * 8 < (1 << 30) else 2
    A = [torch.randint(-2**62, 2**62, (T * L,), dtype=torch.int64, device=dev, generator=g) for _ in range(P)]
    va = [eng.Ciphertext.from_tensor(x, ctx) for x in A]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:16].astype(np.uint64))
    cnt = torch.zeros(P, dtype=torch.int64, device=dev)
    nb = T * L * 8
    row = "N=%-6d L=%-4d %8.1f MB |" % (N, L, nb / 1e6)
    for label, env in forms:
        for k in ("CSGN_DEC_WINDOW", "CSGN_DEC_WINDOW_ALL"): os.environ.pop(k, None)
        os.environ.update(env)
        d = timed(lambda i: key.count_satisfied_async(va[i], cnt.data_ptr() + 8 * i), P)
        row += " %s %6.2f us %.3f |" % (label, d, nb / d / 1e3 / PEAK)
    os.environ.pop("CSGN_DEC_WINDOW", None); os.environ.pop("CSGN_DEC_WINDOW_ALL", None)
    print(row, flush=True)
    del A, va; torch.cuda.empty_cache()
