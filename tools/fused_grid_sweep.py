"""cfg2 / cfg5 fused multiply->decrypt: persistent grid size x rows per item (single calls and a batch of 16).
A fused CTA ends with a global atomic + ticket (fold.cuh); with one item per CTA every item pays that round trip.
    python tools/fused_grid_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import tools.r2_sweep as rs  # noqa: E402
from tools.r2_sweep import setenv, timed, setup, eng

for name, N, D, T1, T2, P in (("cfg2", 1247, 16, 1000, 1000, 16), ("cfg5 300x300", 16383, 64, 300, 300, 12)):
    ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
    nb = T1 * T2 * L * 8
    arr = (eng.handle_array(va), eng.handle_array(vb), eng.handle_array(vo))
    res = []
    for grid in (0, 444, 592, 888, 1184, 1776, 2368, 4736):
        for R in (0, 3, 4, 6, 8, 12):
            kn = {}
            if grid: kn["CSGN_MUL_GRID"] = grid
            if R: kn["CSGN_MUL_R"] = R
            setenv(**kn)
            s = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P)[0]
            b = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=arr), 1, reps=5)[0] / P
            res.append((s, b, grid, R))
    setenv()
    print("# %s: single-call us, batch us per product, grid cap, R   (0 = default)" % name)
    for s, b, grid, R in sorted(res)[:12]:
        print("  single %6.2f (%.3f)  batch %6.2f (%.3f)  grid %-5d R %-2d" % (s, nb / s / 1e3 / rs.PEAK, b, nb / b / 1e3 / rs.PEAK, grid, R))
    d = [r for r in res if r[2] == 0 and r[3] == 0][0]
    print("  default: single %6.2f batch %6.2f" % (d[0], d[1]))
    del va, vb, vo, keep
    import torch; torch.cuda.empty_cache()
