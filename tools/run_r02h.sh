#!/bin/bash
# GPU run r02h (1 GPU): A/B of the 16384-key scatter tile as 1024 threads x 16 keys (main build) against 512 x 32
# (tuning build, -DDNAGPU_SCATTER_512) on c4 / c2 / c5 k=31; the Shift-And scan with a 16 KB stage; partition tests.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_owned.py tests/test_gpu_full_size.py -m gpu -q --maxfail=10 -p no:cacheprovider > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?" > $O/r02h_status.txt
T=$PWD/dna-sequences-pg-extension_b200/libdnagpu_tuning.so
B="python bench.py --cpu-sample 1000000 --no-extract --e2e-steps 2"
for w in c4 c2; do
  $B --workload $w --steps 10 2> $O/r02h_bench_${w}_new.err | grep '^{' > $O/r02h_bench_${w}_new.json; echo "$w new rc=${PIPESTATUS[0]}" >> $O/r02h_status.txt
  DNAGPU_LIB=$T $B --workload $w --steps 10 2> $O/r02h_bench_${w}_old.err | grep '^{' > $O/r02h_bench_${w}_old.json; echo "$w old rc=${PIPESTATUS[0]}" >> $O/r02h_status.txt
done
$B --workload c3 --steps 10 2> $O/r02h_bench_c3_new.err | grep '^{' > $O/r02h_bench_c3_new.json
python tools/ksweep.py --n-bases 1000000000 --seed 5 --ks 13-32 --reps 2 > $O/r02h_ksweep_new.jsonl 2>&1
DNAGPU_LIB=$T python tools/ksweep.py --n-bases 1000000000 --seed 5 --ks 13-32 --reps 2 > $O/r02h_ksweep_old.jsonl 2>&1
cat $O/r02h_status.txt; tail -2 $O/r02h_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02h_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.05})
    except Exception as e:
        print(f, "failed", e)
for f in ("new","old"):
    try:
        rows=[json.loads(l) for l in open(f"gpurun_out/r02h_ksweep_{f}.jsonl") if l.startswith("{")]
        print(f, {r["k"]: round(r["ms"],2) for r in rows})
    except Exception as e:
        print(f, "failed", e)
PY
