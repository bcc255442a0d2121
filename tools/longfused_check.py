"""Fused multiply->decrypt of long blocks: the wide-tile launch rule (csrc/mul.cu, long_fold) against the rule before it
(CSGN_MUL_LONGFOLD=0), single calls on one stream and a batch on the library's lanes (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import tools.r2_sweep as rs  # noqa: E402
from tools.r2_sweep import setenv, timed, setup, eng
import torch

print("# fraction of the measured copy peak: wide tiles (default) | rule before;  single calls, then a batch on the lanes")
for name, N, D, T1, T2, P in (("cfg5 300x300", 16383, 64, 300, 300, 12), ("cfg5 600x300", 16383, 64, 600, 300, 6),
                              ("cfg5 1000x300", 16383, 64, 1000, 300, 4), ("cfg5 1000x1000 (2 GB)", 16383, 64, 1000, 1000, 2),
                              ("N=8191 400x400", 8191, 32, 400, 400, 12), ("N=4096 560x560 (32 units)", 4096, 16, 560, 560, 12),
                              ("N=6000 460x460 (47 units)", 6000, 16, 460, 460, 12), ("N=12000 330x330 (94 units)", 12000, 16, 330, 330, 12),
                              ("N=8191 2000x40 (short rows)", 8191, 32, 2000, 40, 12), ("N=33000 200x200", 33000, 64, 200, 200, 12)):
    ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
    nb = T1 * T2 * L * 8
    arr = (eng.handle_array(va), eng.handle_array(vb), eng.handle_array(vo))
    row = "%-30s %7.1f MB |" % (name, nb / 1e6)
    for kn in ({}, {"CSGN_MUL_LONGFOLD": 0}):
        setenv(**kn)
        s = timed(lambda i: key.mul_count_async(va[i], vb[i], cnt.data_ptr() + 8 * i, out=vo[i]), P)[0]
        b = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=arr), 1, reps=5)[0] / P
        row += " single %7.2f us %.3f  batch %7.2f us %.3f |" % (s, nb / s / 1e3 / rs.PEAK, b, nb / b / 1e3 / rs.PEAK)
    setenv()
    print(row, flush=True)
    del va, vb, vo, keep, arr
    torch.cuda.empty_cache()
