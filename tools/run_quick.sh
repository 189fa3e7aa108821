#!/bin/bash
# quick 1-GPU check of a new build: the partition / owned / full-size tests, then the headline and c2 bench lines
#   tools/run_quick.sh TAG
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-quick}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_owned.py tests/test_gpu_full_size.py -m gpu -q --maxfail=10 -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
B="python bench.py --cpu-sample 1000000 --no-extract --e2e-steps 2"
for w in c4 c2; do
  $B --workload $w --steps 10 2> $O/${TAG}_bench_${w}.err | grep '^{' > $O/${TAG}_bench_${w}.json; echo "$w rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
done
cat $O/${TAG}_status.txt; tail -2 $O/${TAG}_pytest.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.05})
    except Exception as e:
        print(f, "failed", e)
PY
