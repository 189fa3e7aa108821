#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel (first launch) of an .ncu-rep.
usage: tools/ncu_lines.py rep kernel_regex [n]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, line = None, None, None
agg = collections.OrderedDict()
seen_files = set()
stop = False
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file in seen_files:   # second launch of the kernel starts here
            break
        seen_files.add(cur_file)
        continue
    if r[0] in ("Function Name", "Kernel Name"):
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None:
        continue
    if r[0] != "":
        line = (cur_file, int(r[0]), r[1].strip()[:80])
        continue
    try:
        ie = int(r[hdr.index("Instructions Executed")])
        smp = int(r[hdr.index("# Samples")])
    except Exception:
        continue
    a = agg.setdefault(line, [0, 0])
    a[0] += ie
    a[1] += smp
tot = sum(v[0] for v in agg.values())
ts = sum(v[1] for v in agg.values())
print("total warp-instr", tot, "samples", ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100 * v[0] / tot:5.1f}% inst {100 * v[1] / max(ts, 1):5.1f}% smp  {k[0]}:{k[1]}  {k[2]}")
