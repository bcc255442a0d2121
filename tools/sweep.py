"""Tuning sweep for the multiply / decrypt launch knobs (run on the GPU box).

    python tools/sweep.py [cfg2|cfg5|big] > gpurun_out/sweep.txt

Times each knob setting with CUDA events over rotating output buffers (several x the
L2 size), after warm-up.  The knobs are the CSGN_* environment variables the launchers
read on every call; nothing here is part of the product path.
"""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    N, D, T1, T2, nbuf = {"cfg2": (1247, 16, 1000, 1000, 12), "cfg5": (16383, 64, 300, 300, 12),
                          "big": (1247, 16, 5000, 5000, 2), "cfg5big": (16383, 64, 1000, 1000, 3)}[which]
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng.init(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    eng.set_stream(stream.cuda_stream)
    ctx = eng.Context(N, D)
    L = ctx.L
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    a = torch.randint(-2**62, 2**62, (T1 * L,), dtype=torch.int64, device=dev, generator=g)
    b = torch.randint(-2**62, 2**62, (T2 * L,), dtype=torch.int64, device=dev, generator=g)
    outs = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(nbuf)]
    va, vb = eng.Ciphertext.from_tensor(a, ctx), eng.Ciphertext.from_tensor(b, ctx)
    vo = [eng.Ciphertext.from_tensor(o, ctx) for o in outs]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D])
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    out_bytes = T1 * T2 * L * 8
    reps = 5 if which in ("cfg2", "cfg5") else 3

    def timed(fn):
        for i in range(nbuf):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):
            e0.record()
            for _ in range(reps):
                for i in range(nbuf):
                    fn(i)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / (reps * nbuf))
        return best

    def setenv(kv):
        for k in list(os.environ):
            if k.startswith("CSGN_MUL_") or k.startswith("CSGN_DEC_"):
                del os.environ[k]
        for k, v in kv.items():
            os.environ[k] = str(v)

    print("# %s N=%d T1=%d T2=%d out=%.1f MB nbuf=%d" % (which, N, T1, T2, out_bytes / 1e6, nbuf))
    print("# multiply: knobs -> us, GB/s written")
    rows = []
    skip_mul = len(sys.argv) > 2 and sys.argv[2] == "dec"
    tpbs = (320, 480) if L == 20 else (256, 512)
    for U, ips, tpb, st in ([] if skip_mul else itertools.product((1, 2, 4, 8), (16, 32, 64, 128), tpbs, (0, 1, 2))):
        kv = {"CSGN_MUL_U": U, "CSGN_MUL_ITEMS_PER_SM": ips, "CSGN_MUL_TPB": tpb, "CSGN_MUL_STORE": st}
        setenv(kv)
        ms = timed(lambda i: va.mul_into(vb, vo[i]))
        rows.append((ms, kv))
        print("U=%d items/SM=%-3d tpb=%-3d store=%d  %8.2f us  %7.1f GB/s" % (U, ips, tpb, st, ms * 1e3, out_bytes / ms / 1e6), flush=True)
    setenv({})
    ms = timed(lambda i: va.mul_into(vb, vo[i]))
    print("default multiply       %8.2f us  %7.1f GB/s" % (ms * 1e3, out_bytes / ms / 1e6), flush=True)
    rows.sort(key=lambda r: r[0])
    if rows:
        print("# best multiply:", rows[0][1], "%.2f us %.1f GB/s" % (rows[0][0] * 1e3, out_bytes / rows[0][0] / 1e6))
    for R in ([] if skip_mul else (1, 2, 4, 8, 16, 32, 64)):
        kv = dict(rows[0][1])
        kv.pop("CSGN_MUL_ITEMS_PER_SM")
        kv["CSGN_MUL_R"] = R
        setenv(kv)
        ms = timed(lambda i: va.mul_into(vb, vo[i]))
        print("best U/tpb with R=%-2d  %8.2f us  %7.1f GB/s" % (R, ms * 1e3, out_bytes / ms / 1e6), flush=True)
    setenv({"CSGN_MUL_GENERIC": 1})
    ms = timed(lambda i: va.mul_into(vb, vo[i]))
    print("generic u64 kernel     %8.2f us  %7.1f GB/s" % (ms * 1e3, out_bytes / ms / 1e6))

    print("# decrypt: variant x CTAs/SM -> us, GB/s read")
    for var, c in itertools.product((0, 1, 2, 3, 4), (2, 3, 4, 5, 6, 8)):
        setenv({"CSGN_DEC_VARIANT": var, "CSGN_DEC_CTAS_PER_SM": c})
        ms = timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr()))
        print("variant=%d ctas/SM<=%-2d  %8.2f us  %7.1f GB/s" % (var, c, ms * 1e3, out_bytes / ms / 1e6), flush=True)
    setenv({"CSGN_DEC_WIDE": 0})
    ms = timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr()))
    print("ballot kernel (wide off)  %8.2f us  %7.1f GB/s" % (ms * 1e3, out_bytes / ms / 1e6), flush=True)
    setenv({"CSGN_DEC_GENERIC": 1})
    ms = timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr()))
    print("generic warp/block     %8.2f us  %7.1f GB/s" % (ms * 1e3, out_bytes / ms / 1e6))
    setenv({})
    # plain copy for calibration (read + write of the same bytes)
    ms = timed(lambda i: outs[i].copy_(outs[(i + 1) % nbuf]))
    print("# torch copy_ %.2f us  %.1f GB/s (read+write)" % (ms * 1e3, 2 * out_bytes / ms / 1e6))
    ms = timed(lambda i: outs[i].zero_())
    print("# torch zero_ %.2f us  %.1f GB/s (write only)" % (ms * 1e3, out_bytes / ms / 1e6))


if __name__ == "__main__":
    main()
