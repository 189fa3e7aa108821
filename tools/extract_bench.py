#!/usr/bin/env python
"""generate_kmers extraction GB/s on a 1 Gbp synthetic sequence (bytes = 0.25 B/base + 8 B/row)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "dna-sequences-pg-extension_b200")):
    sys.path.insert(0, p)
import torch, dnagpu
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = dnagpu.Context(0, torch_stream=True)
    seq = ctx.synth(n, 5)
    for k in (31, 21, 32, 5):
        rows = seq.kmer_count(k)
        out = torch.empty(rows + 2, dtype=torch.int64, device="cuda")
        for _ in range(3):
            ctx.extract(seq, k, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(stream)
        for _ in range(10):
            ctx.extract(seq, k, out=out)
        e1.record(stream); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        b = 0.25 * n + 8.0 * rows
        print(json.dumps({"k": k, "ms": round(ms, 4), "gbs": round(b / ms / 1e6, 1), "frac_of_6553.9": round(b / ms / 1e6 / 6553.9, 3)}))
        del out
