#!/bin/bash
# GPU run r02w (4 GPUs): the driver's scaling-point command, verbatim, at N = 4 and N = 2 (e2e leg with two shard rings)
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02w}
: > $O/${TAG}_status.txt
for n in 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n bench.py --gpus $n --steps 25 --warmup 5 \
   2> $O/${TAG}_bench_n$n.err | grep '^{' > $O/${TAG}_bench_n$n.json; echo "bench n=$n rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
done
cat $O/${TAG}_status.txt
python - <<PY
import json
for n in (4, 2):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_n%d.json" % n)); print(n, "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2))
    except Exception as e:
        print(n, "failed", e)
PY
