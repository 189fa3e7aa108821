#!/bin/bash
# GPU run r02n (2 GPUs): the code-list collect kernel -- owned/multi tests, the N=8 traffic probe, bench at N=2, C harness.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02n}
timeout 900 python -m pytest tests/test_gpu_owned.py tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
for parts in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$parts \
    tools/probe_collect.py --parts $parts 2> $O/${TAG}_probe_p$parts.err | grep '^{' > $O/${TAG}_probe_p$parts.jsonl; echo "probe $parts rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 \
    bench.py --gpus 2 --steps 10 --warmup 3 --e2e-steps 3 --cpu-sample 1000000 --no-extract --exchange gather \
    2> $O/${TAG}_bench_c4_n2_gather.err | grep '^{' > $O/${TAG}_bench_c4_n2_gather.json; echo "bench rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
timeout 300 ./dna-sequences-pg-extension_b200/dnagpu_bench --bases 3100000000 --k 31 --seed 4 --steps 5 --host --gpus 2 > $O/${TAG}_cbench_n2.json 2> $O/${TAG}_cbench.err
cat $O/${TAG}_status.txt; tail -3 $O/${TAG}_pytest.log; cat $O/${TAG}_probe_p*.jsonl
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_c4_n2_gather.json"))
print("n2 value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms"]/v["launches"],2) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.05})
PY
cat $O/${TAG}_cbench_n2.json
