#!/bin/bash
# GPU run ${TAG} (8 GPUs, or as many as the box has): bench at N = 8 / 4 in the gather form (+ peer at 8 for comparison),
# the C harness on the multi-GPU context, the multi-GPU tests on all GPUs.
cd "$(dirname "$0")/.."
export TAG=${1:-r02g}
O=gpurun_out
NG=$(nvidia-smi -L | wc -l)
nvidia-smi -L > $O/${TAG}_gpus.txt
run_bench() { # n exchange tag [extra args]
  local n=$1 ex=$2 tag=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n \
    bench.py --gpus $n --steps 10 --warmup 3 --e2e-steps 3 --cpu-sample 1000000 --no-extract --exchange $ex "$@" \
    2> $O/${TAG}_bench_c4_n${n}_$tag.err | grep '^{' > $O/${TAG}_bench_c4_n${n}_$tag.json
  echo "bench n=$n $tag rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
}
: > $O/${TAG}_status.txt
run_bench $NG gather gather
[ $NG -ge 8 ] && run_bench 4 gather gather
[ $NG -ge 4 ] && run_bench 2 gather gather
# reads + WHERE (configs[2]) on all GPUs: every rank evaluates the clause on its reads (Shift-And), the rows that pass are routed
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $NG --workload c3 --steps 10 --warmup 3 --e2e-steps 2 --cpu-sample 1000000 \
    2> $O/${TAG}_bench_c3_n${NG}.err | grep '^{' > $O/${TAG}_bench_c3_n${NG}.json; echo "bench c3 n=$NG rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
for g in $NG 4 2 1; do
  [ $g -le $NG ] && timeout 300 ./dna-sequences-pg-extension_b200/dnagpu_bench --bases 3100000000 --k 31 --seed 4 --steps 5 --host --gpus $g \
      > $O/${TAG}_cbench_n$g.json 2>> $O/${TAG}_cbench.err
done
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_owned.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; tail -3 $O/${TAG}_pytest.log
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob("gpurun_out/" + os.environ.get("TAG", "r02g") + "_bench_c[34]_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],2), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms"]/v["launches"],2) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.05})
    except Exception as e:
        print(f, "failed", e)
for f in sorted(glob.glob("gpurun_out/" + os.environ.get("TAG", "r02g") + "_cbench_n*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], d["gpus"], d["gkmer_s"], d["ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
