"""Multi-GPU GROUP BY kmer: one process per GPU, torch.distributed for the plumbing.

The path shards naturally with ONE exchange step:
  1. the sequence is cut into base-range shards, each carrying a (k-1)-base overlap, so every
     k-mer starts in exactly one shard (`shard_of`); reads shard by index with no overlap;
  2. every rank extracts (and filters) its k-mers and buckets them by owner rank,
     owner = dnagpu_owner_of(kmer, world) (libdnagpu: `dnagpu_partition`);
  3. one all-to-all moves each bucket to its owner (NCCL over NVLink/NVSwitch on GPUs);
  4. every rank counts the keys it owns (`dnagpu_count_keys`) -- key sets are disjoint, so the
     three aggregates are plain sums (one 3-element all-reduce) and the grouped result is the
     concatenation of the per-rank tables.

`engine` supplies steps 2 and 4.  `GpuEngine` is the product (libdnagpu on this rank's GPU); the
CPU tests drive the same orchestration over gloo with a stand-in engine.
"""
import numpy as np


def shard_of(n_bases, k, world, rank):
    """Base-range shard of one long sequence -> (first_base, n_starts).

    first_base is a multiple of 32 (a packed-word boundary); the shard's k-mers are those that
    START in [first_base, first_base + n_starts); it needs bases up to first_base+n_starts+k-2."""
    rows = max(0, n_bases - k + 1)
    per = ((rows + world - 1) // world + 31) // 32 * 32
    first = min(rank * per, (rows + 31) // 32 * 32)
    starts = max(0, min(per, rows - first))
    return first, starts


def reads_shard_of(n_reads, world, rank):
    """Index-range shard of a batch of reads -> (first_read, n_reads_local)."""
    first = n_reads * rank // world
    return first, n_reads * (rank + 1) // world - first


class GpuEngine:
    """Steps 2 and 4 on this rank's B200 through the C ABI."""

    def __init__(self, ctx):
        import torch
        self.ctx, self.torch = ctx, torch
        self.send = self.recv = None

    def partition(self, seq, k, world, prefix=None, pattern=None):
        need = seq.kmer_count(k)
        if self.send is None or self.send.numel() < need + 2:
            self.send = self.torch.empty(need + 2, dtype=self.torch.int64, device=f"cuda:{self.ctx.device}")
        return self.ctx.partition(seq, k, world, prefix=prefix, pattern=pattern, out=self.send)

    def recv_buffer(self, n):
        if self.recv is None or self.recv.numel() < n + 2:
            self.recv = self.torch.empty(int(n * 1.05) + 2, dtype=self.torch.int64, device=f"cuda:{self.ctx.device}")
        return self.recv[:n]

    def count_keys(self, keys, k, load_factor=0.0):
        st, _ = self.ctx.count_keys(keys, k, table=False, load_factor=load_factor)
        return st.total, st.distinct, st.unique

    def device(self):
        return self.torch.device("cuda", self.ctx.device)


def count_sharded(engine, seq, k, world, prefix=None, pattern=None, group=None, load_factor=0.0):
    """One pass of the sharded query on this rank -> global (total, distinct, unique)."""
    import torch
    import torch.distributed as dist
    keys, counts = engine.partition(seq, k, world, prefix=prefix, pattern=pattern)
    dev = engine.device()
    send_counts = torch.from_numpy(np.asarray(counts, dtype=np.int64)).to(dev)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    recv_list = [int(x) for x in recv_counts.cpu().tolist()]
    recv = engine.recv_buffer(sum(recv_list))
    dist.all_to_all_single(recv, keys, recv_list, [int(c) for c in counts], group=group)
    t, d, u = engine.count_keys(recv, k, load_factor=load_factor)
    agg = torch.tensor([t, d, u], dtype=torch.int64, device=dev)
    dist.all_reduce(agg, group=group)
    return tuple(int(x) for x in agg.cpu().tolist())
