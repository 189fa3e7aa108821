#!/usr/bin/env python
"""Expected results of the BASELINE.json configurations at their FULL sizes, from the CPU oracle.

    python tests/golden/make_golden_big.py [--only c4] [--threads 8]   ->  tests/golden/big_expected.json

Every entry is the oracle's answer (oracle/ref_cpu.c: ref_count_query_big -- rows produced exactly as the
reference's generate_kmers / starts_with / contains restatement produces them, grouped by the oracle's own hash
aggregate in disjoint hash partitions so that the 3.1 Gbp result fits in memory) for the seeded synthetic input
of include/dnagpu_synth.h: total / distinct / unique as the outer query of README.md:122-130 / test.sql:140-154
computes them, plus the order-independent digest of the grouped (kmer, count) rows (ref_agg_digest).
The GPU tests (tests/test_gpu_full_size.py) and bench.py compare against this file; nothing on the GPU box needs
/root/reference or minutes of CPU time.  Runs in a few minutes on 8 cores; needs ~ 12 GB of RAM for c4."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_cpu as R  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "big_expected.json")
C3 = {"bases": 150, "stride": 5, "prefix": "AC", "pattern": "NNNNNNNNNNNNWSNNNNNNNNNNNNNNNRY"}


def entry(res, dt, **cfg):
    return dict(cfg, total=res.total, distinct=res.distinct, unique=res.unique,
                digest=[int(x) for x in res.digest], oracle_seconds=round(dt, 1))


def seq_case(n, k, seed, threads, faithful=False):
    words = R.synth_seq(seed, n)
    passes = max(1, (n * 8) // (6 << 30) + 1)  # <= 6 GB of keys per pass
    t0 = time.time()
    res = R.count_query_big(words, 1, n, words.size, k, faithful=faithful, passes=passes, threads=threads)
    return entry(res, time.time() - t0, n_bases=n, k=k, seed=seed, repeat_every=8,
                 rows="faithful per-k-mer decode + kmer_make (dna.c:803-825)" if faithful else "two-word window")


def reads_case(n_reads, k, seed, threads, prefix, pattern):
    words = R.synth_reads(seed, n_reads, C3["bases"], C3["stride"])
    t0 = time.time()
    res = R.count_query_big(words, n_reads, C3["bases"], C3["stride"], k, prefix=R.kmer_make(prefix) if prefix else None,
                            pattern=pattern, passes=1, threads=threads)
    return entry(res, time.time() - t0, n_reads=n_reads, bases=C3["bases"], stride=C3["stride"], k=k, seed=seed,
                 repeat_every=8, prefix=prefix, pattern=pattern, rows_tested=n_reads * (C3["bases"] - k + 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 8)
    args = ap.parse_args()
    try:
        with open(OUT) as f:
            out = json.load(f)
    except Exception:
        out = {}
    T = args.threads
    jobs = {
        "c1": lambda: seq_case(10_000, 5, 1, 1, faithful=True),
        "c2": lambda: seq_case(100_000_000, 21, 2, T),
        "c2_faithful": lambda: seq_case(100_000_000, 21, 2, T, faithful=True),
        "c4": lambda: seq_case(3_100_000_000, 31, 4, T),
        "c3": lambda: reads_case(100_000_000, 31, 3, T, C3["prefix"], C3["pattern"]),
        "c3_10m": lambda: reads_case(10_000_000, 31, 3, T, C3["prefix"], C3["pattern"]),
        "c3_10m_prefix": lambda: reads_case(10_000_000, 31, 3, T, C3["prefix"], None),
        "c3_10m_pattern": lambda: reads_case(10_000_000, 31, 3, T, None, C3["pattern"]),
    }
    for k in range(3, 33):
        jobs[f"c5_k{k}"] = (lambda k=k: seq_case(1_000_000_000, k, 5, T))
    for name, fn in jobs.items():
        if args.only and not name.startswith(args.only):
            continue
        t0 = time.time()
        out[name] = fn()
        print(name, out[name], f"{time.time() - t0:.1f}s", flush=True)
        with open(OUT, "w") as f:
            json.dump(out, f, indent=1, sort_keys=True)
            f.write("\n")


if __name__ == "__main__":
    main()
