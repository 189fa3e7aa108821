#!/bin/bash
# 1-GPU A/B of the main build against the tuning build with an environment switch:  tools/run_ab.sh TAG "ENV=1 ..."
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-ab}; SW=${2:-DNAGPU_SCATTER_NOT_PERSISTENT=1}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py tests/test_gpu_owned.py -m gpu -q --maxfail=10 -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
B="python bench.py --cpu-sample 1000000 --no-extract --e2e-steps 2"
for w in c4 c2; do
  $B --workload $w --steps 10 2> $O/${TAG}_bench_${w}_main.err | grep '^{' > $O/${TAG}_bench_${w}_main.json
  env DNAGPU_LIB=$PWD/dna-sequences-pg-extension_b200/libdnagpu_tuning.so $SW $B --workload $w --steps 10 2> $O/${TAG}_bench_${w}_alt.err | grep '^{' > $O/${TAG}_bench_${w}_alt.json
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
