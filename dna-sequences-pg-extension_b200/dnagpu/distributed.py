"""Multi-GPU GROUP BY kmer: one process per GPU, torch.distributed for the plumbing.

The default form (`ShardRing` + `count_sharded_gather`) moves no k-mers at all: the sequence stays sharded by
base range, one shard per GPU with its (k-1)-base overlap, every shard is mapped into every GPU's address space
(CUDA IPC peer memory), and each GPU walks ALL shards -- its own first, the others in ring order -- keeping the
k-mers that dnagpu_owner_of assigns to it (libdnagpu: dnagpu_count with an owner restriction).  NVLink carries
the 2-bit packed bases (0.25 B per start position) instead of 8-byte k-mers, there is no exchange step, and the
only collective is the final 3-element all-reduce.  The forms below it route k-mers to their owners instead.

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


def ring_shards(n_bases, world):
    """k-independent base ranges of one sequence, one per rank -> [(first_base, n_starts)]; first_base is a multiple
    of 32, the ranges tile [0, n_bases), ranks past the end get empty ranges."""
    per = ((n_bases + world - 1) // world + 31) // 32 * 32
    end = (n_bases + 31) // 32 * 32
    out = []
    for r in range(world):
        first = min(r * per, end)
        out.append((first, max(0, min(per, n_bases - first))))
    return out


def shard_words(n_bases, first, starts):
    """Packed words a piece holds: its bases plus the 31-base overlap (any k <= 32), >= 1 zero pad word, even count."""
    held = max(0, min(n_bases, first + starts + 31) - first)
    return ((held + 31) // 32 + 3) & ~1


class ShardRing:
    """Every rank's shard of one sequence in peer-mappable memory, mapped by every rank (one per process).

    Collective constructor: allocates this rank's buffer (dnagpu_peer_alloc), swaps the IPC handles and opens the
    others (dnagpu_peer_open).  Raises on EVERY rank if any rank cannot allocate or map.  `shards` replaces the
    equal base ranges of ring_shards ([(first_base, n_starts)] per rank, first_base a multiple of 32, tiling the
    sequence)."""

    def __init__(self, ctx, world, rank, n_bases, group=None, shards=None):
        import torch.distributed as dist
        self.ctx, self.world, self.rank, self.n_bases, self.group = ctx, world, rank, n_bases, group
        self.shards = list(shards) if shards is not None else ring_shards(n_bases, world)
        if len(self.shards) != world:
            raise ValueError("one base range per rank")
        self.n_words = [shard_words(n_bases, f, s) for f, s in self.shards]
        self.local, self.base, self._seq, handle, err = 0, [], None, None, None
        try:
            self.local, handle = ctx.peer_alloc(8 * self.n_words[rank])
        except Exception as e:  # noqa: BLE001 - reported collectively below
            err = str(e)
        got = [None] * world
        dist.all_gather_object(got, (err, handle), group=group)
        if any(e for e, _ in got):
            if self.local:
                ctx.peer_free(self.local)
            raise RuntimeError("shard ring unavailable: " + "; ".join(e for e, _ in got if e))
        opened = []
        try:
            for r in range(world):
                self.base.append(self.local if r == rank else ctx.peer_open(got[r][1]))
                if r != rank:
                    opened.append(self.base[-1])
        except Exception as e:  # noqa: BLE001
            err = str(e)
        got = [None] * world
        dist.all_gather_object(got, err, group=group)
        if any(got):
            for a in opened:
                ctx.peer_close(a)
            ctx.peer_free(self.local)
            raise RuntimeError("shard ring unavailable: " + "; ".join(e for e in got if e))

    @property
    def my_shard(self):
        return self.shards[self.rank]

    def publish(self):
        """Collective: this rank's shard has been written (ctx.fill_words / ctx.upload_to); wait until all have."""
        import torch.distributed as dist
        self.ctx.synchronize()
        dist.barrier(group=self.group)

    def seq(self):
        """The whole sequence as pieces, this rank's own first and the others in ring order (so that at any
        moment every shard is read by one GPU, not by all of them)."""
        if self._seq is None:
            ring = [(self.rank + i) % self.world for i in range(self.world)]
            self._seq = self.ctx.wrap_pieces([self.base[r] for r in ring], [self.shards[r][0] for r in ring],
                                             [self.shards[r][1] for r in ring], self.n_bases)
        return self._seq

    def close(self):
        import torch.distributed as dist
        if self._seq is not None:
            self._seq.free()
        self.ctx.synchronize()
        dist.barrier(group=self.group)
        for r, a in enumerate(self.base):
            if r != self.rank:
                self.ctx.peer_close(a)
        dist.barrier(group=self.group)
        self.ctx.peer_free(self.local)


def count_sharded_gather(ctx, ring, k, group=None):
    """One pass of the sharded query on this rank -> global (total, distinct, unique).  No k-mer moves: this rank
    counts the k-mers it owns out of the whole sequence (all shards, read through peer memory); key sets are
    disjoint across ranks, so the aggregates add.  The all-reduce also fences the shards for the next upload."""
    import torch
    import torch.distributed as dist
    dev = getattr(ctx, "torch_device", None) or torch.device("cuda", ctx.device)
    st, _ = ctx.count(ring.seq(), k, owner=(ring.world, ring.rank))
    agg = torch.tensor([st.total, st.distinct, st.unique], dtype=torch.int64, device=dev)
    dist.all_reduce(agg, group=group)
    return tuple(int(x) for x in agg.cpu().tolist())


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


class _PairwiseWork:
    """all-to-all over point-to-point sends (backends without a list all-to-all, i.e. gloo on CPU)."""

    def __init__(self, dst_list, src_list, group):
        import torch.distributed as dist
        rank = dist.get_rank(group)
        self.reqs = []
        for r, (d, s_) in enumerate(zip(dst_list, src_list)):
            if r == rank:
                d.copy_(s_)
            else:
                if d.numel():
                    self.reqs.append(dist.irecv(d, src=r, group=group))
                if s_.numel():
                    self.reqs.append(dist.isend(s_.contiguous(), dst=r, group=group))

    def wait(self):
        for q in self.reqs:
            q.wait()


def _all_to_all_list(dst_list, src_list, group):
    import torch.distributed as dist
    if dist.get_backend(group) == "nccl":
        return dist.all_to_all(dst_list, src_list, group=group, async_op=True)
    return _PairwiseWork(dst_list, src_list, group)


def owner_digits(plan, rank):
    """Digits [lo, hi) of partition level 1 that `rank` owns: owner(d) = d * n_parts >> bits1."""
    n = plan.n_digits
    lo = (rank * n + plan.n_parts - 1) // plan.n_parts
    hi = ((rank + 1) * n + plan.n_parts - 1) // plan.n_parts
    return lo, hi


def count_sharded_fused(ctx, seq, k, n_rows_total, world, rank, buffers, prefix=None, pattern=None, group=None,
                        chunks=4):
    """The same query with the owner routing fused into partition level 1 (dnagpu_shuffle_*):
    one scatter pass lays the shard out by hash digit, each owner's digits are one contiguous
    slice of that buffer, the all-to-all ships the slices, and the receiver finishes with
    level 2 + count.

    The exchange is cut into `chunks` digit sub-ranges: all of them are enqueued up front on
    NCCL's stream, and the level 2 + count of sub-range c runs (on the library's stream) while
    sub-range c+1 is still crossing NVLink.  Digits are disjoint across sub-ranges, so the
    aggregates just add.  `buffers` is a dict reused across calls (send / recv tensors)."""
    import os
    import time
    import torch
    import torch.distributed as dist
    dev = getattr(ctx, "torch_device", None) or torch.device("cuda", ctx.device)
    trace = os.environ.get("DNAGPU_TRACE") == "1" and rank == 0 and dev.type == "cuda"
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
    send = buffers["send"]
    digit_counts, kept, side = ctx.shuffle_send(seq, k, plan, send, prefix=prefix, pattern=pattern)
    mark("level1")
    ranges = [owner_digits(plan, r) for r in range(world)]
    lo, hi = ranges[rank]
    n_mine = hi - lo
    # every peer tells me how many keys each of MY digits holds on it (peer-major, digit order)
    counts_dev = torch.from_numpy(digit_counts.astype(np.int64)).to(dev)
    mine_from_all = torch.empty(world * n_mine, dtype=torch.int64, device=dev)
    dist.all_to_all_single(mine_from_all, counts_dev, [n_mine] * world, [b - a for a, b in ranges], group=group)
    pieces = mine_from_all.cpu().numpy().astype(np.uint64).reshape(world, n_mine)
    digit_off = np.concatenate([[0], np.cumsum(digit_counts)]).astype(np.int64)
    n_recv = int(pieces.sum())
    if buffers.get("recv") is None or buffers["recv"].numel() < n_recv + 2:
        buffers["recv"] = torch.empty(int(n_recv * 1.05) + 2, dtype=torch.int64, device=dev)
    mark("counts")
    chunks = max(1, min(chunks, min(b - a for a, b in ranges)))  # every owner has >= 1 digit per sub-range
    jobs, recv_pos = [], 0
    for c in range(chunks):
        # sub-range c of every owner r: digits [a + c*(b-a)//chunks, a + (c+1)*(b-a)//chunks)
        sub = [(a + c * (b - a) // chunks, a + (c + 1) * (b - a) // chunks) for a, b in ranges]
        send_splits = [int(digit_off[y] - digit_off[x]) for x, y in sub]
        mlo, mhi = sub[rank][0] - lo, sub[rank][1] - lo
        my_pieces = np.ascontiguousarray(pieces[:, mlo:mhi])
        recv_splits = [int(v) for v in my_pieces.sum(axis=1)]
        n_c = sum(recv_splits)
        recv_c = buffers["recv"][recv_pos:recv_pos + n_c]
        recv_pos += n_c
        # list form of the all-to-all: the slices of the owners are not adjacent in `send`, no packing copy
        src_list = [send[int(digit_off[x]):int(digit_off[x]) + n] for (x, _), n in zip(sub, send_splits)]
        dst_list, pos = [], 0
        for n in recv_splits:
            dst_list.append(recv_c[pos:pos + n])
            pos += n
        work = _all_to_all_list(dst_list, src_list, group)
        jobs.append((work, recv_c, my_pieces.reshape(-1), mhi - mlo))
    distinct = unique = 0
    for work, recv_c, my_pieces, groups in jobs:
        work.wait()
        st, _ = ctx.shuffle_count(recv_c, my_pieces, groups, plan, k)
        distinct += st.distinct
        unique += st.unique
    mark("exchange+level2+count")
    agg = torch.tensor([kept, distinct, unique, side], dtype=torch.int64, device=dev)
    dist.all_reduce(agg, group=group)
    total, distinct, unique, side_all = (int(x) for x in agg.cpu().tolist())
    mark("all_reduce")
    if trace:
        import sys
        print("trace ms: " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.2f}" for a, b in zip(marks, marks[1:])),
              file=sys.stderr, flush=True)
    return total, distinct + (side_all > 0), unique + (side_all == 1)


class PeerExchange:
    """Receive buffers mapped across the ranks of one node (CUDA IPC), for the exchange that is fused
    into the scatter kernel.  One per process; `capacity_keys` keys per rank."""

    def __init__(self, ctx, world, rank, capacity_keys, group=None):
        """Collective.  Raises on EVERY rank if any rank cannot allocate or map (no IPC / no peer access),
        so that the callers can all fall back to the NCCL exchange together."""
        import torch.distributed as dist
        self.ctx, self.world, self.rank, self.capacity = ctx, world, rank, int(capacity_keys)
        self.local, self.base, handle, err = 0, [], None, None
        try:
            self.local, handle = ctx.peer_alloc(8 * (self.capacity + 2))
        except Exception as e:  # noqa: BLE001 - reported collectively below
            err = str(e)
        got = [None] * world
        dist.all_gather_object(got, (err, handle), group=group)
        if any(e for e, _ in got):
            if self.local:
                ctx.peer_free(self.local)
            raise RuntimeError("peer exchange unavailable: " + "; ".join(e for e, _ in got if e))
        opened = []
        try:
            for r in range(world):
                self.base.append(self.local if r == rank else ctx.peer_open(got[r][1]))
                if r != rank:
                    opened.append(self.base[-1])
        except Exception as e:  # noqa: BLE001
            err = str(e)
        got = [None] * world
        dist.all_gather_object(got, err, group=group)
        if any(got):
            for a in opened:
                ctx.peer_close(a)
            ctx.peer_free(self.local)
            raise RuntimeError("peer exchange unavailable: " + "; ".join(e for e in got if e))

    def close(self):
        import torch.distributed as dist
        dist.barrier()
        for r, a in enumerate(self.base):
            if r != self.rank:
                self.ctx.peer_close(a)
        dist.barrier()
        self.ctx.peer_free(self.local)


def count_sharded_peer(ctx, seq, k, n_rows_total, world, rank, px, prefix=None, pattern=None, group=None):
    """The sharded query with the exchange fused into the level-1 scatter kernel: no all-to-all.

    hist (per-digit counts) -> all-gather of the counts (world x n_digits int64, the only
    collective besides a barrier and the final all-reduce) -> every rank scatters its k-mers
    straight into the owners' receive buffers over NVLink (peer-mapped addresses, one contiguous
    run per (tile, digit)) -> barrier -> level 2 + count of the local receive buffer."""
    import torch
    import torch.distributed as dist
    dev = getattr(ctx, "torch_device", None) or torch.device("cuda", ctx.device)
    # a WHERE clause is evaluated once into a key list (one predicate scan); hist and scatter then read the list,
    # and the partition plan is sized by the rows that passed on all ranks, not by the rows scanned
    listed = ctx.collect(seq, k, prefix=prefix, pattern=pattern) if (prefix is not None or pattern is not None) else None
    if listed is not None:
        passed = torch.tensor([listed.numel()], dtype=torch.int64, device=dev)
        dist.all_reduce(passed, group=group)
        n_rows_total = max(int(passed.item()), 1)
    plan = ctx.shuffle_plan(n_rows_total, world)
    mine = ctx.shuffle_hist_keys(listed, plan) if listed is not None else ctx.shuffle_hist(seq, k, plan)
    counts = torch.empty(world * plan.n_digits, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.from_numpy(mine.astype(np.int64)).to(dev), group=group)
    counts = counts.cpu().numpy().astype(np.uint64).reshape(world, plan.n_digits)
    ranges = [owner_digits(plan, r) for r in range(world)]
    # receive layout of owner o: pieces peer-major, digits of o in order
    dest = np.zeros(plan.n_digits, dtype=np.uint64)
    for o, (a, b) in enumerate(ranges):
        block = counts[:, a:b]                                  # [peer, digit of o]
        before_me = int(block[:rank].sum())
        within = np.concatenate([[0], np.cumsum(block[rank])[:-1]]).astype(np.uint64) if b > a else np.zeros(0, np.uint64)
        if int(block.sum()) > px.capacity:
            raise RuntimeError(f"rank {o} would receive {int(block.sum())} keys, above the peer buffer of {px.capacity}")
        dest[a:b] = np.uint64(px.base[o]) + np.uint64(8) * (np.uint64(before_me) + within)
    if listed is not None:
        kept, side = int(listed.numel()), ctx.shuffle_scatter_keys_to(listed, plan, dest)
    else:
        kept, side = ctx.shuffle_scatter_to(seq, k, plan, dest)
    dist.barrier(group=group)                                   # every rank's stores have landed
    lo, hi = ranges[rank]
    st = ctx.shuffle_count_addr(px.local, counts[:, lo:hi].reshape(-1), hi - lo, plan, k)
    agg = torch.tensor([kept, st.distinct, st.unique, side], dtype=torch.int64, device=dev)
    dist.all_reduce(agg, group=group)                           # also fences the buffers for the next step
    total, distinct, unique, side_all = (int(x) for x in agg.cpu().tolist())
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
