#!/bin/bash
# usage: tools/mgpu_try.sh N "chunks list" [extra bench args]
N=$1; shift; CH=$1; shift
for C in $CH; do
  DNAGPU_TRACE=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N \
    bench.py --gpus $N --steps 3 --warmup 3 --e2e-steps 1 --cpu-sample 1000000 --no-extract --chunks $C "$@" 2> gpurun_out/mgpu_err.log | tail -1 > gpurun_out/mgpu_$C.json
  grep "trace ms" gpurun_out/mgpu_err.log | tail -1
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/mgpu_$C.json")); print("N", d["n_gpus"], "chunks", $C, "value", round(d["value"],2), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],2), d["result"])
except Exception as e:
    print("failed", e); print(open("gpurun_out/mgpu_err.log").read()[-1500:])
PY
done
