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


def owner_digits(plan, rank):
    """Digits [lo, hi) of partition level 1 that `rank` owns: owner(d) = d * n_parts >> bits1."""
    n = plan.n_digits
    lo = (rank * n + plan.n_parts - 1) // plan.n_parts
    hi = ((rank + 1) * n + plan.n_parts - 1) // plan.n_parts
    return lo, hi


def count_sharded_fused(ctx, seq, k, n_rows_total, world, rank, buffers, prefix=None, pattern=None, group=None):
    """The same query with the owner routing fused into partition level 1 (dnagpu_shuffle_*):
    one scatter pass lays the shard out by hash digit, each owner's digits are one contiguous
    slice of that buffer, the all-to-all ships the slices, and the receiver finishes with
    level 2 + count.  `buffers` is a dict reused across calls (send / recv tensors)."""
    import os
    import time
    import torch
    import torch.distributed as dist
    dev = torch.device("cuda", ctx.device)
    trace = os.environ.get("DNAGPU_TRACE") == "1" and rank == 0
    marks = []

    def mark(name):
        if trace:
            torch.cuda.synchronize(dev)
            marks.append((name, time.perf_counter()))
    mark("start")
    plan = ctx.shuffle_plan(n_rows_total, world)
    need = seq.kmer_count(k) + 2
    if buffers.get("send") is None or buffers["send"].numel() < need:
        buffers["send"] = torch.empty(need, dtype=torch.int64, device=dev)
    digit_counts, kept, side = ctx.shuffle_send(seq, k, plan, buffers["send"], prefix=prefix, pattern=pattern)
    mark("level1")
    ranges = [owner_digits(plan, r) for r in range(world)]
    lo, hi = ranges[rank]
    n_mine = hi - lo
    # every peer tells me how many keys each of MY digits holds on it (peer-major, digit order)
    counts_dev = torch.from_numpy(digit_counts.astype(np.int64)).to(dev)
    mine_from_all = torch.empty(world * n_mine, dtype=torch.int64, device=dev)
    dist.all_to_all_single(mine_from_all, counts_dev, [n_mine] * world, [b - a for a, b in ranges], group=group)
    pieces = mine_from_all.cpu().numpy().astype(np.uint64)
    send_splits = [int(digit_counts[a:b].sum()) for a, b in ranges]
    recv_splits = [int(pieces[p * n_mine:(p + 1) * n_mine].sum()) for p in range(world)]
    n_recv = sum(recv_splits)
    if buffers.get("recv") is None or buffers["recv"].numel() < n_recv + 2:
        buffers["recv"] = torch.empty(int(n_recv * 1.05) + 2, dtype=torch.int64, device=dev)
    recv = buffers["recv"][:n_recv]
    mark("counts")
    dist.all_to_all_single(recv, buffers["send"][:sum(send_splits)], recv_splits, send_splits, group=group)
    mark("all_to_all")
    st, _ = ctx.shuffle_count(recv, pieces, n_mine, plan, k)
    mark("level2+count")
    agg = torch.tensor([kept, st.distinct, st.unique, side], dtype=torch.int64, device=dev)
    dist.all_reduce(agg, group=group)
    total, distinct, unique, side_all = (int(x) for x in agg.cpu().tolist())
    mark("all_reduce")
    if trace:
        import sys
        print("trace ms: " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])),
              file=sys.stderr, flush=True)
    return total, distinct + (side_all > 0), unique + (side_all == 1)


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
