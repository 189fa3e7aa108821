#!/usr/bin/env python
"""The sorted k-mer index against the numbers the reference publishes for its SP-GiST index
(test.sql:186-240, ~1 M stored 5-mers: `= 'ATCGC'` 1.3 ms with the index / 41.8 ms seq scan, `^@ 'ACTG'` 4.3 ms /
37.7 ms; the index itself is filled row by row from PL/pgSQL, test.sql:168-179).  Build time and query latency
(wall clock around one C-ABI call, result rows left on the device) for the reference's table size and for 1e9 rows."""
import argparse
import ctypes as C
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-sequences-pg-extension_b200")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="1000000:5,100000000:21,1000000000:31")
    ap.add_argument("--reps", type=int, default=20)
    args = ap.parse_args()
    import torch
    import dnagpu
    from dnagpu import Kmer, _where
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        ctx = dnagpu.Context(0, torch_stream=True)
        for case in args.cases.split(","):
            rows, k = (int(x) for x in case.split(":"))
            g = torch.Generator(device="cuda").manual_seed(1)
            if k <= 16:
                col = torch.randint(0, 4 ** k, (rows,), dtype=torch.int64, device="cuda", generator=g)
            else:                                  # real k-mers of a synthetic sequence (planted repeats)
                seq = ctx.synth(rows + k - 1, 9)
                col = ctx.extract(seq, k)
                seq.free()
            ctx.index_build(col, k).free()         # warm-up (allocator pool)
            ctx.profile(True)
            ctx.profile_reset()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ix = ctx.index_build(col, k)
            torch.cuda.synchronize()
            build_ms = (time.perf_counter() - t0) * 1e3
            kern = {name: ctx.profile_query(name)[0] for name in ("index_keys", "sort_hist", "sort_scatter", "scan")}
            ctx.profile(False)
            passes = (2 * k + 7) // 8
            rec = {"rows": rows, "k": k, "build_ms": build_ms, "build_mrows_s": rows / build_ms / 1e3,
                   "sort_passes": passes, "kernels_ms": kern,
                   "algorithmic_bytes": rows * (8 + 16 + passes * (8 + 32)),
                   "build_gbs": rows * (8 + 16 + passes * (8 + 32)) / build_ms / 1e6}
            probe = int(col[rows // 3].item()) & (2 ** 64 - 1)
            text = str(Kmer(bits=probe, length=k))
            queries = [("equal", text, None), ("prefix4", text[:4], None)]
            if k > 8:
                queries.append(("prefix" + str(k // 2), text[:k // 2], None))
            queries.append(("pattern", None, text[:2] + "N" * (k - 3) + "R"))
            out = torch.empty(rows + 2, dtype=torch.int64, device="cuda")
            n_out = C.c_uint64()
            for name, prefix, pattern in queries:
                if name == "equal":
                    km = Kmer(prefix)
                    call = lambda: ctx.lib.dnagpu_index_equal(ctx.handle, ix.handle, km.bits, km.length, out.data_ptr(),
                                                              out.numel(), C.byref(n_out))
                else:
                    w, _keep = _where(prefix, pattern)
                    call = lambda: ctx.lib.dnagpu_index_search(ctx.handle, ix.handle, C.byref(w), out.data_ptr(),
                                                               out.numel(), C.byref(n_out))
                assert call() == 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(args.reps):
                    assert call() == 0
                ms = (time.perf_counter() - t0) * 1e3 / args.reps
                # the same clause as a sequential scan of the column (dnagpu_filter_keys), for the row count and the time
                kw = dict(prefix=prefix) if pattern is None else dict(pattern=pattern)
                scan = ctx.filter_keys(col, k, **kw)
                assert scan.numel() == n_out.value, (name, scan.numel(), n_out.value)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.filter_keys(col, k, **kw)
                torch.cuda.synchronize()
                rec[name] = {"rows_found": int(n_out.value), "index_ms": ms, "column_scan_ms": (time.perf_counter() - t0) * 1e3}
            print(json.dumps(rec), flush=True)
            ix.free()
            del col, out


if __name__ == "__main__":
    main()
