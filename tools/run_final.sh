#!/bin/bash
# Final check of the tree as it stands (1 GPU): whole -m gpu suite, smoke, the default bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02z}
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_status.txt
timeout 600 python bench.py 2> $O/${TAG}_bench_c4.err | grep '^{' > $O/${TAG}_bench_c4.json; echo "bench rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; tail -2 $O/${TAG}_pytest_gpu.log
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c4.json")); print("value", round(d["value"],2), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), d["e2e"]["steps"], "frac", round(d["roofline"]["frac"],3))
PY
