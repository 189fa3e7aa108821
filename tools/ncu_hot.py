#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples for one kernel of an .ncu-rep.
usage: tools/ncu_hot.py rep kernel_regex [n]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None
data = []
sections = 0
for r in rows:
    if r and r[0] == "Address":
        hdr = r
        sections += 1
        continue
    if hdr and len(r) == len(hdr) and sections == 1:
        data.append(r)
idx = {h: i for i, h in enumerate(hdr)}
tot = sum(int(r[idx["# Samples"]]) for r in data)
inst = sum(int(r[idx["Instructions Executed"]]) for r in data)
print(f"{len(data)} SASS instructions, {tot} samples, {inst} warp-instructions executed")
first_stall = idx["# Samples"] + 1
stall_cols = [h for h in hdr if h.startswith("stall_")]
agg = {h: sum(int(r[idx[h]] or 0) for r in data) for h in stall_cols}
print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]:
    st = {h[6:]: r[idx[h]] for h in stall_cols if r[idx[h]] not in ("0", "")}
    st = dict(sorted(st.items(), key=lambda kv: -int(kv[1]))[:3])
    print(r[idx["# Samples"]].rjust(6), r[idx["Instructions Executed"]].rjust(9), r[idx["Source"]].strip()[:58].ljust(58), st)
