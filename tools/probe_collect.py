#!/usr/bin/env python
"""2-GPU probe of the multi-GPU level-1 kernel (k_collect_owned) under the traffic of a LARGER machine:

    torchrun --nproc-per-node 2 tools/probe_collect.py [--parts 8] [--bases N] [--k 31]

Rank 0 holds 1/parts of the sequence and rank 1 the rest; both count the k-mers owner (parts, rank) out of the
WHOLE sequence.  Rank 0 therefore reads (parts-1)/parts of the bases over NVLink and keeps 1/parts of the
k-mers -- what every GPU of a `parts`-GPU box does -- while rank 1 reads almost everything locally.  Prints the
per-kernel CUDA-event times of both ranks (one JSON line per rank) and checks total = rows / distinct <= total."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dna-sequences-pg-extension_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--bases", type=int, default=3_100_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import dnagpu
    from dnagpu.distributed import ShardRing

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    assert world == 2
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dnagpu.Context(local)
    n, k, G = args.bases, args.k, args.parts
    cut = (n // G + 31) // 32 * 32
    shards = [(0, cut), (cut, n - cut)]
    ring = ShardRing(ctx, world, rank, n, shards=shards)
    first, starts = ring.my_shard
    seq = ctx.synth_range(n, 4, 8, first, starts, 32)
    ctx.fill_words(ring.local, seq, ring.n_words[rank])
    ring.publish()
    whole = ring.seq()
    st, _ = ctx.count(whole, k, owner=(G, rank))          # warm-up
    ctx.profile(True)
    ctx.profile_reset()
    for _ in range(args.steps):
        st, _ = ctx.count(whole, k, owner=(G, rank))
    ctx.synchronize()
    prof = ctx.profile_dump()
    out = {"rank": rank, "parts": G, "peer_fraction": round(1 - starts / n, 4), "rows": n - k + 1, "owned_total": st.total,
           "distinct": st.distinct, "unique": st.unique,
           "kernels_ms": {name: round(v["ms"] / max(1, v["launches"]), 4) for name, v in prof.items()}}
    assert abs(st.total * G - (n - k + 1)) < 0.01 * (n - k + 1), "an owner keeps ~ 1/parts of the rows"
    assert st.unique <= st.distinct <= st.total
    for r in range(world):
        if r == rank:
            print(json.dumps(out), flush=True)
        dist.barrier(device_ids=[local])
    ring.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
