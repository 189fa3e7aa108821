#!/usr/bin/env python
"""BASELINE.json configs[4]: k-sweep 3..32 over a 1 Gbp synthetic sequence (seed 5).
Prints one JSON line per k with the per-kernel CUDA-event times; optionally checks the three
aggregates against the CPU oracle on a prefix (--check-bases)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-sequences-pg-extension_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-bases", type=int, default=1_000_000_000)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--ks", default="3-32")
    ap.add_argument("--method", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check-bases", type=int, default=0)
    args = ap.parse_args()
    import dnagpu
    lo, _, hi = args.ks.partition("-")
    ks = list(range(int(lo), int(hi or lo) + 1))
    ctx = dnagpu.Context(0)
    seq = ctx.synth(args.n_bases, args.seed)
    small = None
    if args.check_bases:
        from oracle import ref_cpu as R
        words = R.synth_seq(args.seed, args.check_bases)
        small = ctx.upload(dnagpu.Dna.from_words(words, args.check_bases))
    for k in ks:
        st, _ = ctx.count(seq, k, method=args.method)  # warm-up
        ctx.profile(True)
        ctx.profile_reset()
        for _ in range(args.reps):
            st, _ = ctx.count(seq, k, method=args.method)
        prof = ctx.profile_dump()
        ctx.profile(False)
        ms = sum(v["ms"] for v in prof.values()) / args.reps
        line = {"k": k, "n_bases": args.n_bases, "ms": ms, "gkmer_s": st.total / ms / 1e6,
                "total": st.total, "distinct": st.distinct, "unique": st.unique,
                "kernels": {n: v["ms"] / args.reps for n, v in prof.items()}}
        if small is not None:
            want = R.count_query(words, 1, args.check_bases, words.size, k, faithful=False, want_rows=False,
                                 threads=8)
            got, _ = ctx.count(small, k, method=args.method)
            line["parity_prefix"] = (got.total, got.distinct, got.unique) == want.stats
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
