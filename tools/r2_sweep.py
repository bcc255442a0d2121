"""Round-2 timing sweeps (run on the GPU box; CUDA events, rotating buffers larger than L2).

    python tools/r2_sweep.py fused      # cfg2: multiply, decrypt, fused multiply->decrypt, count-only; grid caps
    python tools/r2_sweep.py chain      # chain shapes (10^6 rows x 25 / 125 blocks): tiled knobs vs the flat kernel
    python tools/r2_sweep.py shapes     # N in {191 .. 33000} x {mul, decrypt, fused, permute, add}: fraction of peak
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CSGN_TUNING", "1")
from csgn_b200 import engine as eng  # noqa: E402

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
eng.init(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
gen = torch.Generator(device=dev)
gen.manual_seed(1)
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0

KNOBS = [k for k in os.environ if k.startswith("CSGN_MUL_") or k.startswith("CSGN_DEC_") or k.startswith("CSGN_PERM_")]


def setenv(**kv):
    # R2_EXTRA="CSGN_MUL_GRID=592,CSGN_X=1": knobs applied under every measurement of a run (what-if on a whole table)
    for item in os.environ.get("R2_EXTRA", "").split(","):
        if "=" in item:
            k_, v_ = item.split("=", 1)
            kv.setdefault(k_, v_)
    for k in list(os.environ):
        if (k.startswith("CSGN_MUL_") or k.startswith("CSGN_DEC_") or k.startswith("CSGN_PERM_")) and k not in kv:
            del os.environ[k]
    for k, v in kv.items():
        os.environ[k] = str(v)


def timed(fn, n_items, reps=5, rounds=3):
    """median over `rounds` of (time of reps passes over n_items calls) / (reps*n_items), in us"""
    for i in range(n_items):
        fn(i)
    torch.cuda.synchronize()
    res = []
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            for i in range(n_items):
                fn(i)
        e1.record()
        torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 1e3 / (reps * n_items))
    return float(np.median(res)), float(min(res))


def rand_words(n):
    return torch.randint(-2**62, 2**62, (n,), dtype=torch.int64, device=dev, generator=gen)


def setup(N, D, T1, T2, P):
    ctx = eng.Context(N, D)
    L = ctx.L
    A = [rand_words(T1 * L) for _ in range(P)]
    B = [rand_words(T2 * L) for _ in range(P)]
    O = [torch.empty(T1 * T2 * L, dtype=torch.int64, device=dev) for _ in range(P)]
    va = [eng.Ciphertext.from_tensor(x, ctx) for x in A]
    vb = [eng.Ciphertext.from_tensor(x, ctx) for x in B]
    vo = [eng.Ciphertext.from_tensor(x, ctx) for x in O]
    key = eng.SecretKey(ctx, np.random.default_rng(7).permutation(N)[:D].astype(np.uint64))
    cnt = torch.zeros(P, dtype=torch.int64, device=dev)
    return ctx, L, va, vb, vo, key, cnt, (A, B, O)


def line(label, us, bytes_, extra=""):
    med, best = us if isinstance(us, tuple) else (us, us)
    gbs = bytes_ / med / 1e3
    print("  %-58s %9.2f us (min %8.2f) %7.0f GB/s  %.3f of peak %s" % (label, med, best, gbs, gbs / PEAK, extra), flush=True)


def sec_fused():
    N, D, T1, T2, P = 1247, 16, 1000, 1000, 16
    ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
    nb = T1 * T2 * L * 8
    print("# cfg2 1000x1000, %d rotating products of %.0f MB; peak %.0f GB/s" % (P, nb / 1e6, PEAK))
    setenv()
    line("multiply (single calls)", timed(lambda i: va[i].mul_into(vb[i], vo[i]), P), nb)
    line("decrypt (single calls)", timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr() + 8 * i), P), nb)
    line("fused multiply->decrypt (single calls)", timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P), nb)
    line("fused count-only (single calls)", timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i), P), nb)
    arr = (eng.handle_array(va), eng.handle_array(vb), eng.handle_array(vo))
    t = timed(lambda i: eng.mul_into_batch(None, None, None, arrays=arr), 1, reps=5)
    line("multiply batch of %d (per product)" % P, (t[0] / P, t[1] / P), nb)
    t = timed(lambda i: key.count_satisfied_batch_async(None, cnt.data_ptr(), array=arr[2]), 1, reps=5)
    line("decrypt batch of %d (per product)" % P, (t[0] / P, t[1] / P), nb)
    t = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=arr), 1, reps=5)
    line("fused batch of %d (per product)" % P, (t[0] / P, t[1] / P), nb)
    t = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=(arr[0], arr[1], None)), 1, reps=5)
    line("fused count-only batch of %d (per product)" % P, (t[0] / P, t[1] / P), nb)
    combos = [dict(CSGN_MUL_ALIGN=0), dict(CSGN_MUL_R=3), dict(CSGN_MUL_R=4), dict(CSGN_MUL_R=6), dict(CSGN_MUL_R=8), dict(CSGN_MUL_R=12),
              dict(CSGN_MUL_R=16), dict(CSGN_MUL_R=4, CSGN_MUL_U=1), dict(CSGN_MUL_R=8, CSGN_MUL_U=1), dict(CSGN_MUL_R=4, CSGN_MUL_TPB=384),
              dict(CSGN_MUL_R=4, CSGN_MUL_TPB=256), dict(CSGN_MUL_R=8, CSGN_MUL_TPB=256), dict(CSGN_MUL_GRID=148 * 4), dict(CSGN_MUL_GRID=148 * 8), dict(CSGN_MUL_GRID=148 * 16),
              dict(CSGN_MUL_ITEMS_PER_SM=16), dict(CSGN_MUL_ITEMS_PER_SM=64), dict(CSGN_MUL_U=4), dict(CSGN_MUL_U=1),
              dict(CSGN_MUL_R=8), dict(CSGN_MUL_R=16), dict(CSGN_MUL_R=32), dict(CSGN_MUL_TPB=320), dict(CSGN_MUL_TPB=256)]
    for g_ in (148 * 3, 148 * 4, 148 * 5, 148 * 6, 148 * 8):
        for r_ in (4, 8, 16, 32):
            combos.append(dict(CSGN_MUL_GRID=g_, CSGN_MUL_R=r_))
    for u_, r_ in ((4, 8), (4, 16), (1, 16), (1, 32), (1, 64)):
        combos.append(dict(CSGN_MUL_U=u_, CSGN_MUL_R=r_))
        combos.append(dict(CSGN_MUL_U=u_, CSGN_MUL_R=r_, CSGN_MUL_GRID=148 * 4))
    for knobs in combos:
        setenv(**knobs)
        t = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=arr), 1, reps=5)
        line("fused batch, %s" % knobs, (t[0] / P, t[1] / P), nb)
        t = timed(lambda i: eng.mul_into_batch(None, None, None, arrays=arr), 1, reps=5)
        line("multiply batch, %s" % knobs, (t[0] / P, t[1] / P), nb)
    setenv()


def sec_chain():
    for name, Td, P in (("chain25", 25, 2), ("chain125", 125, 1)):
        N, D, T1 = 1247, 16, 1000000
        ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, Td, P)
        nb = T1 * Td * L * 8
        print("# %s: 10^6 rows x %d blocks = %.1f GB per product, %d buffers" % (name, Td, nb / 1e9, P))
        reps = 3 if Td == 25 else 2
        z = keep[2]
        def zero(i):
            z[i].zero_()
        setenv()
        line("torch zero_ (write-only reference)", timed(zero, P, reps=reps), nb)
        line("tiled default", timed(lambda i: va[i].mul_into(vb[i], vo[i]), P, reps=reps), nb)
        for knobs in (dict(CSGN_MUL_R=4), dict(CSGN_MUL_R=8), dict(CSGN_MUL_R=16), dict(CSGN_MUL_R=32),
                      dict(CSGN_MUL_U=2), dict(CSGN_MUL_U=4, CSGN_MUL_R=8), dict(CSGN_MUL_U=4, CSGN_MUL_R=16),
                      dict(CSGN_MUL_U=2, CSGN_MUL_R=8), dict(CSGN_MUL_ITEMS_PER_SM=32), dict(CSGN_MUL_ITEMS_PER_SM=512),
                      dict(CSGN_MUL_TPB=256), dict(CSGN_MUL_TPB=512), dict(CSGN_MUL_PF_CTAS_PER_SM=0),
                      dict(CSGN_MUL_PF_CTAS_PER_SM=16)):
            setenv(**knobs)
            line("tiled %s" % knobs, timed(lambda i: va[i].mul_into(vb[i], vo[i]), P, reps=reps), nb)
        for knobs in (dict(), dict(CSGN_MUL_FLAT_U=1), dict(CSGN_MUL_FLAT_U=2), dict(CSGN_MUL_FLAT_U=8),
                      dict(CSGN_MUL_FLAT_CTAS_PER_SM=2), dict(CSGN_MUL_FLAT_CTAS_PER_SM=3), dict(CSGN_MUL_FLAT_CTAS_PER_SM=6),
                      dict(CSGN_MUL_FLAT_CTAS_PER_SM=8), dict(CSGN_MUL_FLAT_CTAS_PER_SM=16), dict(CSGN_MUL_FLAT_CTAS_PER_SM=64),
                      dict(CSGN_MUL_TPB=256), dict(CSGN_MUL_TPB=512), dict(CSGN_MUL_TPB=512, CSGN_MUL_FLAT_CTAS_PER_SM=2),
                      dict(CSGN_MUL_FLAT_U=2, CSGN_MUL_FLAT_CTAS_PER_SM=8), dict(CSGN_MUL_FLAT_U=8, CSGN_MUL_FLAT_CTAS_PER_SM=2)):
            setenv(CSGN_MUL_FLAT=1, **knobs)
            line("flat %s" % knobs, timed(lambda i: va[i].mul_into(vb[i], vo[i]), P, reps=reps), nb)
        setenv()
        line("decrypt of the product", timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr() + 8 * i), P, reps=reps), nb)
        line("fused tiled", timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P, reps=reps), nb)
        setenv(CSGN_MUL_FLAT=1)
        line("fused flat", timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P, reps=reps), nb)
        line("fused flat count-only", timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i), P, reps=reps), nb)
        setenv()
        del va, vb, vo, keep, z
        torch.cuda.empty_cache()


def sec_shapes():
    print("# fraction of the measured copy peak (%.0f GB/s); mul/decrypt/fused: 8L bytes per block, permute/add: 16L" % PEAK)
    rows = []
    for N, D, T in ((191, 4, 3000), (1247, 16, 1000), (2048, 16, 800), (4097, 16, 560), (8191, 32, 400), (16383, 64, 300),
                    (33000, 64, 200)):
        P = max(2, min(16, int(2.6e9 // (T * T * eng.words_per_block(N) * 8))))
        ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T, T, P)
        nb = T * T * L * 8
        setenv()
        r = {"N": N, "L": L, "blocks": T * T, "MB": nb / 1e6}
        r["mul"] = timed(lambda i: va[i].mul_into(vb[i], vo[i]), P)[0]
        r["decrypt"] = timed(lambda i: key.count_satisfied_async(vo[i], cnt.data_ptr() + 8 * i), P)[0]
        r["fused"] = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P)[0]
        perm = eng.Permutation(ctx, np.random.default_rng(3).permutation(N).astype(np.uint64))
        r["permute"] = timed(lambda i: vo[i].permute_into(perm, vo[(i + 1) % P]), P)[0] if P >= 2 else None
        def add(i):
            s_ = vo[i] + vo[(i + 1) % P]
            del s_
        r["add"] = timed(add, P)[0]
        frac = {k: (nb * (2 if k in ("permute",) else 4 if k == "add" else 1)) / r[k] / 1e3 / PEAK for k in ("mul", "decrypt", "fused", "permute", "add")}
        rows.append((r, frac))
        print("  N=%-6d L=%-4d %8d blocks %7.1f MB | " % (N, L, T * T, nb / 1e6) +
              " | ".join("%s %8.2f us %.3f" % (k, r[k], frac[k]) for k in ("mul", "decrypt", "fused", "permute", "add")), flush=True)
        del va, vb, vo, keep
        torch.cuda.empty_cache()
    print(json.dumps([{**r, **{"frac_" + k: v for k, v in f.items()}} for r, f in rows]))




def sec_rsel():
    """rows per item (R) for the plain and the fused multiply over the shapes the launcher has to choose for"""
    shapes = [("cfg2 1000x1000", 1247, 16, 1000, 1000, 16), ("N=1247 10000x1000 (1.6 GB)", 1247, 16, 10000, 1000, 2),
              ("N=1247 1000x10000 (1.6 GB)", 1247, 16, 1000, 10000, 2),
              ("cfg5 300x300", 16383, 64, 300, 300, 12), ("cfg5 1000x1000 (2 GB)", 16383, 64, 1000, 1000, 2),
              ("cfg5 2000x2000 (8.2 GB)", 16383, 64, 2000, 2000, 1), ("chain25 (4 GB)", 1247, 16, 1000000, 25, 2),
              ("chain125 (20 GB)", 1247, 16, 1000000, 125, 1), ("N=191 3000x3000", 191, 4, 3000, 3000, 8),
              ("N=4097 560x560", 4097, 16, 560, 560, 12), ("N=33000 200x200", 33000, 64, 200, 200, 12)]
    for name, N, D, T1, T2, P in shapes:
        ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
        nb = T1 * T2 * L * 8
        reps = 5 if nb < 1e9 else 2
        print("# %s: %.1f MB per product, %d buffers" % (name, nb / 1e6, P))
        for R in (0, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64):
            for U in ((0,) if R == 0 else (0, 1, 2, 4)):
                kn = {}
                if R:
                    kn["CSGN_MUL_R"] = R
                if U:
                    kn["CSGN_MUL_U"] = U
                setenv(**kn)
                m = timed(lambda i: va[i].mul_into(vb[i], vo[i]), P, reps=reps)
                f = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P, reps=reps)
                print("  R=%-3s U=%-2s  multiply %9.2f us %.3f | fused %9.2f us %.3f" %
                      (R or "dflt", U or "d", m[0], nb / m[0] / 1e3 / PEAK, f[0], nb / f[0] / 1e3 / PEAK), flush=True)
        setenv()
        del va, vb, vo, keep
        torch.cuda.empty_cache()


def sec_longfused():
    """fused multiply->decrypt of long blocks (shared-memory fold, persistent grid): rows per item x units per thread x
    CTAs per SM of the persistent grid, single stream (the way bench.py's other_workloads times it)"""
    shapes = [("cfg5 300x300", 16383, 64, 300, 300, 12), ("N=8191 400x400", 8191, 32, 400, 400, 12),
              ("N=4097 560x560", 4097, 16, 560, 560, 12), ("cfg5 1000x300", 16383, 64, 1000, 300, 4)]
    for name, N, D, T1, T2, P in shapes:
        ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
        nb = T1 * T2 * L * 8
        print("# %s: %.1f MB per product, %d buffers" % (name, nb / 1e6, P))
        setenv()
        f = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P)
        print("  default                      fused %9.2f us %.3f" % (f[0], nb / f[0] / 1e3 / PEAK), flush=True)
        for U in (2, 4):
            for R in (2, 3, 4, 5, 6, 7, 8, 10):
                row = "  U=%d R=%-2d |" % (U, R)
                for cps in (4, 8, 16, 1000):
                    setenv(CSGN_MUL_R=R, CSGN_MUL_U=U, CSGN_MUL_FOLD_CTAS_PER_SM=cps)
                    f = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P)
                    row += " %4d/SM %7.2f us %.3f |" % (cps, f[0], nb / f[0] / 1e3 / PEAK)
                print(row, flush=True)
        setenv()
        del va, vb, vo, keep
        torch.cuda.empty_cache()


if __name__ == "__main__" and len(sys.argv) >= 1 and os.path.basename(sys.argv[0]) == "r2_sweep.py":
    for which in (sys.argv[1:] or ["fused", "chain", "shapes"]):
        {"fused": sec_fused, "chain": sec_chain, "shapes": sec_shapes, "rsel": sec_rsel, "longfused": sec_longfused}[which]()
