"""Block-range sharding of a ciphertext over one process per GPU.

The hot path shards without any data-path collective: every output block
of a multiply depends on one block of each operand (reference src/Ciphertext.cpp:159),
decrypt is a per-block predicate folded by XOR (src/SecretKey.cpp:131-140), permute is
per block.  So

  * a ciphertext is split by CONTIGUOUS block range, rank g owning csgn_shard_range(g);
  * for a*b the LEFT operand is the sharded one and b is replicated: rank g then owns
    output blocks [first*T2, (first+count)*T2) -- contiguous and in the reference's own
    i-major order, so chains (a*b)*d with replicated d stay shard-local;
  * the only exchange is decrypt's: the per-rank satisfied-block counts are summed (the parity
    of the sum is the XOR of the parities).  On GPUs the fold kernel does that exchange itself:
    its last CTA stores the count into every rank's mailbox over NVLink and the launch that
    closes a batch collects the sums (engine.PeerComm, csrc/peer.cuh) -- no collective library
    on the data path.  torch.distributed is only the rendezvous that carries the 64-byte
    mailbox handles (connect_peers).  allreduce_counts (one all-reduce) remains for gloo/CPU
    test doubles and as the baseline the fused path is measured against.

The compute (engine.Ciphertext) is passed in, so the same logic runs under gloo on CPU
in tests/ with the oracle standing in for the kernels.
"""
import torch
import torch.distributed as dist

from . import _native


def shard_range(n_blocks, rank, world):
    """(first, count) of the contiguous block range rank owns -- the C ABI's csgn_shard_range."""
    import ctypes
    lib = _native.load()
    first, count = ctypes.c_uint64(), ctypes.c_uint64()
    _native.check(lib.csgn_shard_range(int(n_blocks), int(rank), int(world), ctypes.byref(first), ctypes.byref(count)))
    return first.value, count.value


def world():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_counts(counts, group=None):
    """Sum per-rank satisfied-block counts in place (int64 tensor, any length).  One collective
    for a whole batch of decrypts; a no-op in a single-process run."""
    if world() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def connect_peers(group=None):
    """engine.PeerComm over the ranks of `group`: exchanges the mailbox handles with one all_gather."""
    from . import engine
    w, r = world(), rank()

    def allgather_bytes(b):
        nccl = dist.get_backend(group) == "nccl"
        dev = torch.device("cuda", torch.cuda.current_device()) if nccl else torch.device("cpu")
        mine = torch.tensor(list(b), dtype=torch.uint8, device=dev)
        out = [torch.empty_like(mine) for _ in range(w)]
        dist.all_gather(out, mine, group=group)
        return [bytes(t.cpu().tolist()) for t in out]

    return engine.PeerComm(r, w, allgather_bytes if w > 1 else None)


def parity(counts):
    """decrypt = XOR over blocks = parity of the satisfied-block count."""
    return counts & 1


class ShardedCiphertext:
    """This rank's contiguous slice of a ciphertext of `global_blocks` blocks."""

    def __init__(self, local, first, global_blocks):
        self.local = local                  # engine.Ciphertext (or a test double) with this rank's blocks
        self.first = int(first)             # global index of the first local block
        self.global_blocks = int(global_blocks)

    @classmethod
    def scatter_from_host(cls, words, ctx, make_ciphertext):
        """Every rank holds (or can generate) the full operand; each uploads only its range."""
        n = len(words) // ctx.L
        first, count = shard_range(n, rank(), world())
        return cls(make_ciphertext(words[first * ctx.L:(first + count) * ctx.L], ctx), first, n)

    @property
    def count(self):
        return self.local.n_blocks

    def mul_replicated(self, other):
        """self * other with `other` replicated on every rank: shard-local, no communication.
        The result is again a contiguous range, [first*T2, (first+count)*T2)."""
        t2 = other.n_blocks
        return ShardedCiphertext(self.local * other, self.first * t2, self.global_blocks * t2)

    def permute(self, perm):
        """applyPermutation on every block: shard-local."""
        return ShardedCiphertext(self.local.applyPermutation(perm), self.first, self.global_blocks)

    def decrypt(self, key, counts_out=None, device=None, comm=None):
        """Local fold and the one-word exchange.  Returns the plaintext bit (int).
        With a PeerComm the fold kernel does the exchange itself; otherwise one all-reduce follows."""
        if comm is not None:
            return comm.decrypt(key, self.local)[0]
        if counts_out is None:
            if device is None:
                nccl = dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl"
                device = "cuda" if nccl else "cpu"
            c = torch.tensor([key.count_satisfied(self.local)], dtype=torch.int64, device=device)
        else:
            c = counts_out                          # caller's tensor: element 0 receives the local count, then the sum
            c.view(-1)[0] = key.count_satisfied(self.local)
        allreduce_counts(c)
        return int(parity(c)[0].item())
