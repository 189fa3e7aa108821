#!/bin/bash
# GPU run r02e (1 GPU): full GPU test suite on the current build, then A/B of the two kernel variants of this step
# (Shift-And step on the FMA pipe; bucket count looking back from the placing thread) against the tuning build that
# keeps the earlier forms, then ncu of the new forms.
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?" > $O/r02e_status.txt
T=$PWD/dna-sequences-pg-extension_b200/libdnagpu_tuning.so
B="python bench.py --cpu-sample 1000000 --no-extract --e2e-steps 2"
for w in c4 c3 c2; do
  $B --workload $w --steps 10 | grep '^{' > $O/r02e_bench_${w}_new.json 2> $O/r02e_bench_${w}_new.err; echo "$w new rc=${PIPESTATUS[0]}" >> $O/r02e_status.txt
  DNAGPU_LIB=$T $B --workload $w --steps 10 | grep '^{' > $O/r02e_bench_${w}_old.json 2> $O/r02e_bench_${w}_old.err; echo "$w old rc=${PIPESTATUS[0]}" >> $O/r02e_status.txt
done
F="python bench.py --workload c3 --n-bases 1500000000 --steps 1 --warmup 3 --e2e-steps 1 --cpu-sample 1000000"
$F > /dev/null 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_filter_sa" -s 3 -c 1 -f -o $O/r02e_c3_filter_sa $F > $O/r02e_ncu_f.log 2>&1; echo "ncu sa rc=$?" >> $O/r02e_status.txt
K="python tools/ksweep.py --n-bases 1000000000 --seed 5 --ks 31 --reps 1"
$K > $O/r02e_ksweep_k31.json 2>&1 && timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_count_buckets_bins" -s 1 -c 1 \
    -f -o $O/r02e_1gbp_count_bins $K > $O/r02e_ncu_count.log 2>&1; echo "ncu count rc=$?" >> $O/r02e_status.txt
cat $O/r02e_status.txt; tail -3 $O/r02e_pytest.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02e_bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "value", round(d["value"],1), "ms", round(d["ms_per_step"],3), {k:round(v["ms"]/v["launches"],3) for k,v in d["kernels"].items() if v["ms"]/v["launches"]>0.05})
    except Exception as e:
        print(f, "failed", e)
PY
