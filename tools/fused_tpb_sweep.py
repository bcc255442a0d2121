"""cfg2 fused multiply->decrypt: CTA size x rows per item x units per thread, lane-aligned and shared-memory fold
(batch of 16 rotating products, CUDA events).   python tools/fused_tpb_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.argv = [sys.argv[0]]
import tools.r2_sweep as rs  # noqa: E402  (initialises the engine)
from tools.r2_sweep import setenv, timed, line, setup, eng

N, D, T1, T2, P = 1247, 16, 1000, 1000, 16
ctx, L, va, vb, vo, key, cnt, keep = setup(N, D, T1, T2, P)
nb = T1 * T2 * L * 8
arr = (eng.handle_array(va), eng.handle_array(vb), eng.handle_array(vo))
res = []
for align in (1, 0):
    for tpb in (128, 192, 256, 320, 384, 512):
        for R in (4, 6, 8, 12, 16):
            for U in (1, 2, 4):
                if R * U > 64:
                    continue
                setenv(CSGN_MUL_ALIGN=align, CSGN_MUL_TPB=tpb, CSGN_MUL_R=R, CSGN_MUL_U=U)
                t = timed(lambda i: eng.mul_count_batch_async(key, None, None, cnt.data_ptr(), arrays=arr), 1, reps=5)
                res.append((t[0] / P, align, tpb, R, U))
setenv()
res.sort()
for us, align, tpb, R, U in res[:25]:
    print("  %6.2f us  align=%d tpb<=%d R=%d U=%d   %.3f of peak" % (us, align, tpb, R, U, nb / us / 1e3 / rs.PEAK))
print("  ...")
for us, align, tpb, R, U in res[-3:]:
    print("  %6.2f us  align=%d tpb<=%d R=%d U=%d" % (us, align, tpb, R, U))
