#!/usr/bin/env python
"""Predicate scan over a materialised kmer column -- the only operations the reference
publishes timings for (test.sql:191-261: `=`, `^@ 'ACTG'`, `'MRKYN' @>` over ~1 M stored 5-mers, seq scan
37-42 ms on one core = ~25 Mkmer/s).  Same predicates through dnagpu_filter_keys on a column resident in HBM."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-sequences-pg-extension_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_000_000_000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    import numpy as np
    import torch
    import dnagpu
    from oracle import ref_cpu as R
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = dnagpu.Context(0, torch_stream=True)
        g = torch.Generator(device="cuda").manual_seed(1)
        col = torch.randint(0, 4 ** args.k, (args.rows,), dtype=torch.int64, device="cuda", generator=g)
        cases = [("kmer = 'ATCGC'", dict(prefix="ATCGC")), ("kmer ^@ 'ACTG'", dict(prefix="ACTG")),
                 ("'MRKYN' @> kmer", dict(pattern="MRKYN"))]
        sample = col[:200_000].cpu().numpy().view(np.uint64)
        for name, kw in cases:
            out = ctx.filter_keys(col, args.k, **kw)  # warm-up, also gives the row count
            # parity on a sample against the oracle's operators
            got = ctx.filter_keys(col[:200_000], args.k, **kw).cpu().numpy().view(np.uint64)
            keep = np.ones(sample.size, dtype=bool)
            if "prefix" in kw:
                pb, pl = R.kmer_make(kw["prefix"])
                keep &= (sample & np.uint64((1 << (2 * pl)) - 1)) == np.uint64(pb)
            want = np.array([x for x in sample[keep] if "pattern" not in kw or R.contains(kw["pattern"], int(x), args.k)],
                            dtype=np.uint64)
            assert np.array_equal(got, want), name
            # one C-ABI call per scan into a preallocated result buffer (count per tile, scan, ordered write)
            import ctypes as C
            from dnagpu import _where
            w, _keep = _where(kw.get("prefix"), kw.get("pattern"))
            res = torch.empty(out.numel() + 2, dtype=torch.int64, device="cuda")
            n_out = C.c_uint64()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record(stream)
            for _ in range(args.reps):
                rc = ctx.lib.dnagpu_filter_keys(ctx.handle, col.data_ptr(), col.numel(), args.k, C.byref(w),
                                                res.data_ptr(), res.numel(), C.byref(n_out))
                assert rc == 0 and n_out.value == out.numel()
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            ctx.profile(True); ctx.profile_reset()
            out = ctx.filter_keys(col, args.k, **kw)
            prof = {k2: round(v["ms"], 3) for k2, v in ctx.profile_dump().items()}
            ctx.profile(False)
            print(json.dumps({"predicate": name, "rows": args.rows, "k": args.k, "matches": int(out.numel()),
                              "ms": ms, "gkmer_s": args.rows / ms / 1e6,
                              "hbm_gbs": (8.0 * args.rows * 2 + 8.0 * out.numel()) / ms / 1e6, "kernels_ms": prof,
                              "note": "two passes over the column (count per tile, ordered write) + rows out",
                              "reference_seq_scan_mkmer_s": {"kmer = 'ATCGC'": 23.9, "kmer ^@ 'ACTG'": 26.5,
                                                             "'MRKYN' @> kmer": 26.7}[name]}), flush=True)


if __name__ == "__main__":
    main()
