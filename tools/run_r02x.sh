#!/bin/bash
# GPU run r02x (8 GPUs): the driver's scaling-point command, verbatim, at N = 8
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02x}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 25 --warmup 5 \
   2> $O/${TAG}_bench_n8.err | grep '^{' > $O/${TAG}_bench_n8.json; echo "bench n=8 rc=${PIPESTATUS[0]}" > $O/${TAG}_status.txt
cat $O/${TAG}_status.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n8.json")); print(8, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2))
PY
