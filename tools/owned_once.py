#!/usr/bin/env python
"""One owner-restricted count on one GPU (what every GPU of a `parts`-GPU box runs on the whole sequence), for ncu:
    ncu -k regex:k_collect_owned ... python tools/owned_once.py --parts 8 --bases 1000000000"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dna-sequences-pg-extension_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--part", type=int, default=0)
    ap.add_argument("--bases", type=int, default=1_000_000_000)
    ap.add_argument("--k", type=int, default=31)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    import dnagpu
    ctx = dnagpu.Context(0)
    seq = ctx.synth(args.bases, 5, 8)
    ctx.profile(True)
    for _ in range(args.reps):
        ctx.profile_reset()
        st, _ = ctx.count(seq, args.k, owner=(args.parts, args.part))
        ctx.synchronize()
    prof = ctx.profile_dump()
    print(json.dumps({"parts": args.parts, "bases": args.bases, "k": args.k, "owned_total": st.total, "distinct": st.distinct,
                      "unique": st.unique, "kernels_ms": {n: round(v["ms"] / max(1, v["launches"]), 4) for n, v in prof.items()}}))


if __name__ == "__main__":
    main()
