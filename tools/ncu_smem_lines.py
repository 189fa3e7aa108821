#!/usr/bin/env python
"""Shared-memory wavefronts per SOURCE LINE of one kernel of an .ncu-rep (--set full --import-source on):
ideal / actual / excessive L1 wavefronts (bank conflicts) of every line that touches shared memory.
usage: tools/ncu_smem_lines.py rep kernel_regex"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}",
                      "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
cur_file, hdr, line, seen = None, None, None, set()
agg = collections.OrderedDict()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        if cur_file in seen:
            break
        seen.add(cur_file)
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or r[0] in ("Function Name", "Kernel Name"):
        continue
    if r[0] != "":
        line = (cur_file, int(r[0]), r[1].strip()[:90])
        continue
    try:
        w = int(r[hdr.index("L1 Wavefronts Shared")] or 0)
        wi = int(r[hdr.index("L1 Wavefronts Shared Ideal")] or 0)
        we = int(r[hdr.index("L1 Wavefronts Shared Excessive")] or 0)
        ie = int(r[hdr.index("Instructions Executed")] or 0)
    except Exception:
        continue
    a = agg.setdefault(line, [0, 0, 0, 0])
    a[0] += w; a[1] += wi; a[2] += we; a[3] += ie
tot = sum(v[0] for v in agg.values())
print(f"kernel {kern}: {tot} shared-memory wavefronts, {sum(v[2] for v in agg.values())} of them excessive (bank conflicts), "
      f"{sum(v[3] for v in agg.values())} warp instructions")
print("share  wavefronts      ideal  excessive  line")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if v[0] == 0:
        continue
    print(f"{100 * v[0] / max(tot, 1):5.1f}% {v[0]:11d} {v[1]:10d} {v[2]:10d}  {k[0]}:{k[1]}  {k[2]}")
