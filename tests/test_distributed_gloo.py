"""The N > 1 orchestration (dnagpu/distributed.py) on CPU: two gloo ranks, the same shard /
owner-routing / all-to-all / all-reduce code the GPU run uses, with a stand-in engine (oracle
extraction, libdnagpu's host-callable dnagpu_owner_of for routing) in place of the CUDA kernels."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


class OracleEngine:
    """Test double for GpuEngine: steps 2 and 4 done by the CPU oracle."""

    def __init__(self, n_bases, seed):
        self.n_bases, self.seed = n_bases, seed

    def device(self):
        return torch.device("cpu")

    def partition(self, seq, k, world, prefix=None, pattern=None):
        from oracle import ref_cpu as R
        import dnagpu
        first, starts = seq
        n_local = min(self.n_bases - first, starts + k - 1)
        words = R.synth_seq(self.seed, self.n_bases, first_word=first // 32, n_words=(n_local + 31) // 32 + 1)
        rows = R.generate_kmers(words, n_local, k, window=True)[:starts]
        owners = np.array([dnagpu.owner_of(int(x), world) for x in rows], dtype=np.int64)
        order = np.argsort(owners, kind="stable")
        counts = np.bincount(owners, minlength=world).astype(np.uint64)
        return torch.from_numpy(rows[order].view(np.int64).copy()), counts

    def recv_buffer(self, n):
        return torch.empty(n, dtype=torch.int64)

    def count_keys(self, keys, k, load_factor=0.0):
        u, c = np.unique(keys.numpy().view(np.uint64), return_counts=True)
        return int(c.sum()), int(u.size), int((c == 1).sum())


class OracleShuffleCtx:
    """Test double for dnagpu.Context in count_sharded_fused: the plan comes from libdnagpu's host-callable
    dnagpu_shuffle_plan_make, extraction / counting from the CPU oracle and numpy."""
    MUL = 0xD6E8FEB86659FD93

    def __init__(self, n_bases, seed):
        self.n_bases, self.seed, self.device = n_bases, seed, 0
        self.torch_device = torch.device("cpu")

    def shuffle_plan(self, n_rows_total, n_parts):
        import ctypes as C
        from dnagpu import _lib
        plan = _lib.ShufflePlan()
        assert _lib.load().dnagpu_shuffle_plan_make(n_rows_total, n_parts, C.byref(plan)) == 0
        return plan

    def shuffle_send(self, seq, k, plan, out, prefix=None, pattern=None):
        from oracle import ref_cpu as R
        first, starts = seq.shard
        n_local = min(self.n_bases - first, starts + k - 1)
        words = R.synth_seq(self.seed, self.n_bases, first_word=first // 32, n_words=(n_local + 31) // 32 + 1)
        rows = R.generate_kmers(words, n_local, k, window=True)[:starts]
        side = int((rows == np.uint64(2**64 - 1)).sum())
        rows = rows[rows != np.uint64(2**64 - 1)]
        digits = np.array([((int(x) * self.MUL) & (2**64 - 1)) >> (64 - plan.bits1) for x in rows], dtype=np.int64)
        order = np.argsort(digits, kind="stable")
        out[:rows.size] = torch.from_numpy(rows[order].view(np.int64).copy())
        return np.bincount(digits, minlength=plan.n_digits).astype(np.uint64), rows.size + side, side

    def shuffle_count(self, keys, piece_counts, n_groups, plan, k, table=False):
        assert int(np.sum(piece_counts)) == keys.numel() and len(piece_counts) % n_groups == 0

        class St:
            pass
        u, c = np.unique(keys.numpy().view(np.uint64), return_counts=True)
        st = St()
        st.total, st.distinct, st.unique = int(c.sum()), int(u.size), int((c == 1).sum())
        return st, None


class OraclePeerCtx(OracleShuffleCtx):
    """Test double for the peer exchange (count_sharded_peer + PeerExchange): receive buffers are POSIX shared
    memory blocks mapped by every rank, addressed like device memory (a made-up base per owner + byte offset),
    so the layout arithmetic, the barriers and the WHERE path run exactly as on the GPUs."""

    def __init__(self, n_bases, seed, rank):
        super().__init__(n_bases, seed)
        self.rank, self.blocks, self.views = rank, [], {}

    # ---- peer memory ----
    def peer_alloc(self, n_bytes):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=max(int(n_bytes), 16))
        base = (self.rank + 1) << 44
        self.blocks.append(shm)
        self.views[base] = np.ndarray(shm.size // 8, dtype=np.uint64, buffer=shm.buf)
        self.own = shm
        return base, (shm.name, base)

    def peer_open(self, handle):
        from multiprocessing import shared_memory
        name, base = handle
        shm = shared_memory.SharedMemory(name=name)
        self.blocks.append(shm)
        self.views[base] = np.ndarray(shm.size // 8, dtype=np.uint64, buffer=shm.buf)
        return base

    def peer_close(self, addr):
        self.views.pop(addr, None)

    def peer_free(self, addr):
        self.views.clear()
        for shm in self.blocks:
            shm.close()
        self.own.unlink()

    def _store(self, addr, rows):
        base = int(addr) >> 44 << 44
        off = (int(addr) - base) // 8
        self.views[base][off:off + rows.size] = rows

    # ---- what the kernels do ----
    def _rows(self, seq, k, prefix=None, pattern=None):
        from oracle import ref_cpu as R
        first, starts = seq.shard
        n_local = min(self.n_bases - first, starts + k - 1)
        words = R.synth_seq(self.seed, self.n_bases, first_word=first // 32, n_words=(n_local + 31) // 32 + 1)
        rows = R.generate_kmers(words, n_local, k, window=True)[:starts]
        if prefix is not None:
            pb, pl = R.kmer_make(prefix)
            rows = rows[(rows & np.uint64((1 << (2 * pl)) - 1)) == np.uint64(pb)]
        if pattern is not None:
            rows = np.array([x for x in rows if R.contains(pattern, int(x), k)], dtype=np.uint64)
        return rows

    def _digits(self, rows, plan):
        return np.array([((int(x) * self.MUL) & (2**64 - 1)) >> (64 - plan.bits1) for x in rows], dtype=np.int64)

    def collect(self, seq, k, prefix=None, pattern=None):
        return torch.from_numpy(self._rows(seq, k, prefix, pattern).view(np.int64).copy())

    def shuffle_hist(self, seq, k, plan):
        return self.shuffle_hist_keys(self.collect(seq, k), plan)

    def shuffle_hist_keys(self, keys, plan):
        rows = keys.numpy().view(np.uint64)
        rows = rows[rows != np.uint64(2**64 - 1)]
        return np.bincount(self._digits(rows, plan), minlength=plan.n_digits).astype(np.uint64)

    def shuffle_scatter_to(self, seq, k, plan, digit_dest):
        keys = self.collect(seq, k)
        return keys.numel(), self.shuffle_scatter_keys_to(keys, plan, digit_dest)

    def shuffle_scatter_keys_to(self, keys, plan, digit_dest):
        rows = keys.numpy().view(np.uint64)
        side = int((rows == np.uint64(2**64 - 1)).sum())
        rows = rows[rows != np.uint64(2**64 - 1)]
        digits = self._digits(rows, plan)
        for d in np.unique(digits):
            self._store(digit_dest[d], rows[digits == d])
        return side

    def shuffle_count_addr(self, addr, piece_counts, n_groups, plan, k):
        n = int(np.sum(piece_counts))
        keys = torch.from_numpy(self.views[int(addr)][:n].view(np.int64).copy())
        return self.shuffle_count(keys, piece_counts, n_groups, plan, k)[0]


class OracleRingCtx(OraclePeerCtx):
    """Test double for ShardRing + count_sharded_gather: shards live in shared-memory blocks every rank maps; the
    owner-restricted count rebuilds the sequence from the pieces it was handed (in the order it was handed them),
    extracts with the CPU oracle and keeps what libdnagpu's host-callable dnagpu_owner_of assigns to the rank."""

    def synchronize(self):
        pass

    def fill_shard(self, addr, first, starts, n_words):
        from oracle import ref_cpu as R
        held = max(0, min(self.n_bases, first + starts + 31) - first)
        words = np.zeros(n_words, dtype=np.uint64)
        w = R.synth_seq(self.seed, self.n_bases, first_word=first // 32, n_words=(held + 31) // 32)
        if held % 32:  # bases past the piece's reach exist in the sequence but not in the piece
            w[-1] &= np.uint64((1 << (2 * (held % 32))) - 1)
        words[:w.size] = w
        self._store(addr, words)

    def wrap_pieces(self, addrs, first_bases, n_starts, n_bases_total, keep=None):
        class Pieces:
            pass
        p = Pieces()
        p.addrs, p.first, p.starts, p.n_bases = list(addrs), list(first_bases), list(n_starts), n_bases_total
        p.free = lambda: None
        return p

    def count(self, seq, k, owner=None, **kw):
        from oracle import ref_cpu as R
        import dnagpu
        G, me = owner
        rows_total = max(0, seq.n_bases - k + 1)
        keys = []
        for addr, first, starts in zip(seq.addrs, seq.first, seq.starts):
            starts = max(0, min(starts, rows_total - first))
            if not starts:
                continue
            held = min(seq.n_bases, first + starts + k - 1) - first
            base = int(addr) >> 44 << 44
            words = self.views[base][: (held + 31) // 32 + 1].copy()
            keys.append(R.generate_kmers(words, held, k, window=True)[:starts])
        keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
        mine = np.array([dnagpu.owner_of(int(x), G) == me for x in keys], dtype=bool)
        u, c = np.unique(keys[mine], return_counts=True)

        class St:
            pass
        st = St()
        st.total, st.distinct, st.unique = int(c.sum()), int(u.size), int((c == 1).sum())
        return st, None


def _worker_gather(rank, world, port, n_bases, k, seed, out, shards=None):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dnagpu.distributed import ShardRing, count_sharded_gather
    ctx = OracleRingCtx(n_bases, seed, rank)
    ring = ShardRing(ctx, world, rank, n_bases, shards=shards)
    first, starts = ring.my_shard
    ctx.fill_shard(ring.local, first, starts, ring.n_words[rank])
    ring.publish()
    res = [count_sharded_gather(ctx, ring, k) for _ in range(2)]
    ring.close()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bases,k,world", [(30_000, 31, 2), (20_011, 21, 3), (100, 32, 2), (40, 5, 4), (4_097, 13, 2)])
def test_gather_form_equals_single_rank(n_bases, k, world):
    """ShardRing + count_sharded_gather (bench.py's default N > 1 path): k-independent base-range shards with the
    31-base overlap in memory every rank maps, pieces walked in ring order, owner-restricted count, all-reduce."""
    from oracle import ref_cpu as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (n_bases + k + world) % 150
    procs = [ctx.Process(target=_worker_gather, args=(r, world, port, n_bases, k, 31, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words = R.synth_seq(31, n_bases)
    want = R.count_query(words, 1, n_bases, words.size, k, faithful=False)
    assert [tuple(g) for g in got] == [want.stats, want.stats]


def test_gather_form_with_uneven_shards():
    """ShardRing(shards=...): base ranges of the caller's choosing (tools/probe_collect.py gives rank 0 an eighth of the
    sequence), one of them empty."""
    from oracle import ref_cpu as R
    n_bases, k, world = 10_000, 21, 3
    shards = [(0, 1_248), (1_248, n_bases - 1_248), (n_bases // 32 * 32 + 32, 0)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_gather, args=(r, world, 29391, n_bases, k, 31, q, shards)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words = R.synth_seq(31, n_bases)
    want = R.count_query(words, 1, n_bases, words.size, k, faithful=False)
    assert [tuple(g) for g in got] == [want.stats, want.stats]


def test_ring_shards_tile_the_sequence_and_hold_their_overlap():
    from dnagpu.distributed import ring_shards, shard_words
    for n, world in [(3_100_000_000, 8), (1000, 3), (40, 4), (33, 2), (1, 8), (100_000_007, 7)]:
        pos = 0
        for first, starts in ring_shards(n, world):
            assert first % 32 == 0
            if starts:
                assert first == pos
            pos += starts
            held = max(0, min(n, first + starts + 31) - first)
            assert shard_words(n, first, starts) >= (held + 31) // 32 + 1 and shard_words(n, first, starts) % 2 == 0
        assert pos == n


class FakeSeq:
    def __init__(self, shard, k):
        self.shard, self.k = shard, k

    def kmer_count(self, k):
        return self.shard[1]


def _worker_fused(rank, world, port, n_bases, k, seed, chunks, out):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dnagpu.distributed import count_sharded_fused, shard_of
    ctx = OracleShuffleCtx(n_bases, seed)
    seq = FakeSeq(shard_of(n_bases, k, world, rank), k)
    res = count_sharded_fused(ctx, seq, k, n_bases - k + 1, world, rank, {}, chunks=chunks)
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def _worker_peer(rank, world, port, n_bases, k, seed, where, out):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dnagpu.distributed import PeerExchange, count_sharded_peer, shard_of
    ctx = OraclePeerCtx(n_bases, seed, rank)
    seq = FakeSeq(shard_of(n_bases, k, world, rank), k)
    px = PeerExchange(ctx, world, rank, n_bases)
    res = [count_sharded_peer(ctx, seq, k, n_bases - k + 1, world, rank, px, **where) for _ in range(2)]  # buffers reused
    px.close()
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


def _worker(rank, world, port, n_bases, k, seed, out):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dnagpu.distributed import count_sharded, shard_of
    engine = OracleEngine(n_bases, seed)
    shard = shard_of(n_bases, k, world, rank)
    res = count_sharded(engine, shard, k, world)
    if rank == 0:
        out.put(res)
    dist.destroy_process_group()


@pytest.mark.parametrize("n_bases,k", [(20_000, 31), (4_097, 5), (64, 32)])
def test_two_rank_exchange_equals_single_rank(n_bases, k):
    from oracle import ref_cpu as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (n_bases + k) % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_bases, k, 21, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words = R.synth_seq(21, n_bases)
    want = R.count_query(words, 1, n_bases, words.size, k, faithful=False)
    assert tuple(got) == want.stats


@pytest.mark.parametrize("n_bases,k,chunks", [(30_000, 31, 1), (30_000, 21, 3), (100, 32, 2)])
def test_two_rank_fused_exchange_equals_single_rank(n_bases, k, chunks):
    """count_sharded_fused: level-1 digit layout, owner digit ranges, peer-major pieces, chunked all-to-all."""
    from oracle import ref_cpu as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (n_bases + k + chunks) % 90
    procs = [ctx.Process(target=_worker_fused, args=(r, 2, port, n_bases, k, 23, chunks, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words = R.synth_seq(23, n_bases)
    want = R.count_query(words, 1, n_bases, words.size, k, faithful=False)
    assert tuple(got) == want.stats


@pytest.mark.parametrize("n_bases,k,world,where", [(30_000, 31, 2, {}), (20_000, 21, 3, {}), (100, 32, 2, {}),
                                                   (30_000, 9, 2, dict(prefix="AC")),
                                                   (30_000, 9, 2, dict(prefix="G", pattern="NNNNWNNNS"))])
def test_peer_exchange_equals_single_rank(n_bases, k, world, where):
    """count_sharded_peer (bench.py's default N > 1 path): per-digit counts all-gathered, every rank stores its
    runs at the computed offsets inside the owners' (here: shared-memory) buffers, barrier, count; with a WHERE
    clause the key list is collected first and the plan is sized by the rows that passed on all ranks."""
    from oracle import ref_cpu as R
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29750 + (n_bases + k + world + len(where)) % 100
    procs = [ctx.Process(target=_worker_peer, args=(r, world, port, n_bases, k, 29, where, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    words = R.synth_seq(29, n_bases)
    want = R.count_query(words, 1, n_bases, words.size, k, faithful=False, pattern=where.get("pattern"),
                         prefix=R.kmer_make(where["prefix"]) if "prefix" in where else None)
    assert [tuple(g) for g in got] == [want.stats, want.stats]


def test_owner_digit_ranges_agree_with_the_library():
    import ctypes as C
    from dnagpu import _lib
    from dnagpu.distributed import owner_digits
    lib = _lib.load()
    for n, parts in [(3_100_000_000, 8), (3_100_000_000, 2), (1_000_000, 3), (5_000_000_000, 7), (100, 4)]:
        plan = _lib.ShufflePlan()
        assert lib.dnagpu_shuffle_plan_make(n, parts, C.byref(plan)) == 0
        assert plan.n_digits == 1 << plan.bits1 and plan.bits2 >= 1 and plan.n_digits >= parts
        seen = 0
        for r in range(parts):
            lo, hi = owner_digits(plan, r)
            assert lo == seen and hi > lo
            assert all(lib.dnagpu_shuffle_owner(C.byref(plan), d) == r for d in (lo, hi - 1))
            seen = hi
        assert seen == plan.n_digits


def test_shards_tile_the_sequence_exactly_once():
    from dnagpu.distributed import reads_shard_of, shard_of
    for n, k, world in [(1000, 31, 2), (3_100_000_000, 31, 8), (33, 32, 4), (10, 31, 3), (100_000_007, 21, 7)]:
        rows = max(0, n - k + 1)
        pos = 0
        for r in range(world):
            first, starts = shard_of(n, k, world, r)
            assert first % 32 == 0
            if starts:
                assert first == pos
            pos += starts
            assert first + starts <= max(rows, first)
        assert pos == rows
    for n, world in [(10, 3), (100_000_000, 8), (1, 4)]:
        tot = 0
        for r in range(world):
            first, cnt = reads_shard_of(n, world, r)
            assert first == tot
            tot += cnt
        assert tot == n
