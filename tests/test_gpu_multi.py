"""More than one GPU (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`):
the C-ABI multi-GPU context (dnagpu_create_multi: one process, peer access) and the one-process-per-GPU form
(ShardRing + count_sharded_gather over CUDA IPC) against the CPU oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import dnagpu
from conftest import PKG, ROOT
from oracle import ref_cpu as R

pytestmark = pytest.mark.gpu


def _n_gpus():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def test_multi_context_on_one_device_is_a_plain_context():
    ctx = dnagpu.Context(devices=[0])
    assert ctx.lib.dnagpu_device_count(ctx.handle) == 1
    n, k = 300_000, 21
    words = R.synth_seq(3, n)
    want = R.count_query(words, 1, n, words.size, k, faithful=False)
    st, table = ctx.count_kmers(dnagpu.Dna.from_words(words, n), k)
    kmers, counts = table.sorted()
    assert (st.total, st.distinct, st.unique) == want.stats
    assert np.array_equal(kmers, want.kmers) and np.array_equal(counts, want.counts)
    ctx.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("k", [13, 31, 32])
def test_create_multi_count_kmers_matches_oracle(k):
    G = min(_n_gpus(), 8)
    ctx = dnagpu.Context(devices=list(range(G)))
    assert ctx.lib.dnagpu_device_count(ctx.handle) == G
    n, seed = 40_000_000, 5
    words = R.synth_seq(seed, n)
    want = R.count_query_big(words, 1, n, words.size, k, threads=8)
    dna = dnagpu.Dna.from_words(words, n)
    for _ in range(2):                       # the shard buffers are reused
        st, table = ctx.count_kmers(dna, k)
        assert (st.total, st.distinct, st.unique) == want.stats
        assert table.rows == want.distinct
        kmers, counts = table.fetch()
        assert np.array_equal(R.pairs_digest(kmers, counts), want.digest)
        a, b = table.fetch(table.rows // 3, 1000)  # a range that may straddle two GPUs' rows
        assert np.array_equal(a, kmers[table.rows // 3: table.rows // 3 + 1000])
        assert np.array_equal(b, counts[table.rows // 3: table.rows // 3 + 1000])
        with pytest.raises(dnagpu.DnaError):
            table.device_pointers()
        table.free()
    # a small input and a WHERE clause take the single-GPU path of the same context
    small = dnagpu.Dna("ACGTACGTACGTAG")
    assert dnagpu.kmer_stats(small, 8, ctx=ctx) == (7, 5, 3)          # test.sql:107-119
    st, _ = ctx.count_kmers(dna, k, prefix="AC", table=False)
    w = R.count_query_big(words, 1, n, words.size, k, prefix=R.kmer_make("AC"), threads=8)
    assert (st.total, st.distinct, st.unique) == w.stats
    ctx.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
def test_c_harness_multi_gpu_host_entry():
    G = min(_n_gpus(), 8)
    exe = os.path.join(PKG, "dnagpu_bench")
    n, k, seed = 50_000_000, 31, 4
    out = subprocess.run([exe, "--bases", str(n), "--k", str(k), "--seed", str(seed), "--steps", "2", "--host",
                          "--gpus", str(G)], check=True, capture_output=True, text=True).stdout
    import json
    line = json.loads(out)
    words = R.synth_seq(seed, n)
    want = R.count_query_big(words, 1, n, words.size, k, threads=8)
    assert line["gpus"] == G and (line["total"], line["distinct"], line["unique"]) == want.stats


_RING_WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
import torch, torch.distributed as dist
import dnagpu
from dnagpu.distributed import ShardRing, count_sharded_gather
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
n, k, seed = {n}, {k}, {seed}
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = dnagpu.Context(rank, torch_stream=True)
    ring = ShardRing(ctx, world, rank, n)
    first, starts = ring.my_shard
    s = ctx.synth_range(n, seed, 8, first, starts, 32)
    ctx.fill_words(ring.local, s, ring.n_words[rank])
    s.free()
    ring.publish()
    res = [count_sharded_gather(ctx, ring, k) for _ in range(2)]
    ring.close()
    ctx.close()
if rank == 0:
    print(json.dumps(res))
dist.destroy_process_group()
"""


@pytest.mark.skipif(_n_gpus() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("k", [21, 32])
def test_shard_ring_gather_one_process_per_gpu_matches_oracle(k, tmp_path):
    G = min(_n_gpus(), 8)
    n, seed = 60_000_000, 5
    script = tmp_path / "ring_worker.py"
    script.write_text(_RING_WORKER.format(root=ROOT, pkg=PKG, n=n, k=k, seed=seed))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={G}",
                          "--master-addr", "127.0.0.1", "--master-port", str(29300 + k), str(script)],
                         check=True, capture_output=True, text=True, timeout=600).stdout
    import json
    got = json.loads(out.strip().splitlines()[-1])
    words = R.synth_seq(seed, n)
    want = R.count_query_big(words, 1, n, words.size, k, threads=8)
    assert [tuple(g) for g in got] == [want.stats, want.stats]
