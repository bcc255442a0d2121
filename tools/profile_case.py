"""Small fixed workload for ncu (run plain first, then under ncu -- see profiles/README.md).

    python tools/profile_case.py [cfg2|cfg5|big] [n_iter]

Launches, per iteration: multiply (rotating output buffers), decrypt of an older
product, then one permute and one concat at the end.  No timing is reported here:
numbers taken under a profiler are never bench values.  Also one batched encryption of T1*T2 bits.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from csgn_b200 import engine as eng  # noqa: E402


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    N, D, T1, T2, nbuf = {"cfg2": (1247, 16, 1000, 1000, 4), "cfg5": (16383, 64, 300, 300, 4),
                          "big": (1247, 16, 5000, 5000, 2)}[which]
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
    perm = eng.Permutation(ctx, np.random.default_rng(8).permutation(N))
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    for i in range(iters):
        va.mul_into(vb, vo[i % nbuf])
        key.count_satisfied_async(vo[(i + 1) % nbuf], cnt.data_ptr())
    # the batch entry points: two products and two folds spread over the library's lanes
    eng.mul_into_batch([va, va], [vb, vb], [vo[0], vo[1 % nbuf]])
    cnt2 = torch.zeros(2, dtype=torch.int64, device=dev)
    key.count_satisfied_batch_async([vo[0], vo[1 % nbuf]], cnt2.data_ptr())
    # the sharded-decrypt path at world size 1: fold + publish + collect in the one kernel (csrc/peer.cuh)
    comm = eng.PeerComm(0, 1)
    tot = torch.zeros(2, dtype=torch.int64, device=dev)
    comm.push(key, vo[0])
    comm.push(key, vo[1 % nbuf], 2, tot.data_ptr())
    vo[0].permute_into(perm, vo[1])
    # the fused multiply -> decrypt (product written and folded in one launch), and its count-only form
    key.mul_count_async(va, vb, cnt.data_ptr(), out=vo[2 % nbuf])
    key.mul_count_async(va, vb, cnt.data_ptr())
    fresh = key.encrypt_batch(np.random.default_rng(9).integers(0, 2, size=T1 * T2).astype(np.uint8), seed=1)
    s = va + vb
    torch.cuda.synchronize()
    print("profile_case", which, "done; launches:", eng.launch_count(), "sum blocks", s.n_blocks)


if __name__ == "__main__":
    main()
