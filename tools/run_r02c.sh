#!/bin/bash
# GPU run r02c (1 GPU): the other workloads' bench lines (c3 in both WHERE forms, c2, c5 sweep, c1), the reference arm,
# then ncu: launch list + DRAM bytes of the headline workload, --set full of the partition kernels (1 Gbp, k = 31)
# and of the two predicate-scan kernels (10 M reads).  Every ncu command runs only after the plain command exited 0.
cd "$(dirname "$0")/.."
O=gpurun_out
B="python bench.py --cpu-sample 2000000"
$B --workload c3 --steps 5 > $O/r02c_bench_c3.json 2> $O/r02c_bench_c3.err; echo "c3 rc=$?" > $O/r02c_status.txt
$B --workload c3 --steps 5 --planes > $O/r02c_bench_c3_planes.json 2> $O/r02c_bench_c3_planes.err; echo "c3 planes rc=$?" >> $O/r02c_status.txt
$B --workload c2 --steps 20 > $O/r02c_bench_c2.json 2> $O/r02c_bench_c2.err; echo "c2 rc=$?" >> $O/r02c_status.txt
$B --workload c5 --steps 3 > $O/r02c_bench_c5_sweep.json 2> $O/r02c_bench_c5.err; echo "c5 rc=$?" >> $O/r02c_status.txt
$B --workload c1 --steps 50 > $O/r02c_bench_c1.json 2> $O/r02c_bench_c1.err; echo "c1 rc=$?" >> $O/r02c_status.txt
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02c_bench_reference.json 2> $O/r02c_bench_reference.err; echo "ref rc=$?" >> $O/r02c_status.txt
python bench.py --impl reference --workload c1 --steps 20 --warmup 2 > $O/r02c_bench_reference_c1.json 2>> $O/r02c_bench_reference.err
# ---- ncu ----
L="python bench.py --steps 1 --warmup 3 --e2e-steps 1 --cpu-sample 1000000 --no-extract"
$L > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
    --log-file $O/r02c_c4_launches_dram.csv $L > $O/r02c_ncu_c4.log 2>&1; echo "ncu launches rc=$?" >> $O/r02c_status.txt
# the two scatter kernels at the headline fan-outs (2048 / 1024) on 100 Mbp: the tuning build forces the bucket bits
export DNAGPU_LIB=$PWD/dna-sequences-pg-extension_b200/libdnagpu_tuning.so
K="python tools/ksweep.py --n-bases 100000000 --seed 2 --ks 31 --reps 1"
DNAGPU_BUCKET_BITS=21 $K > $O/r02c_ksweep_fan21.json 2>&1 && DNAGPU_BUCKET_BITS=21 timeout 900 ncu --set full --import-source on --clock-control none \
    -k regex:"k_part_scatter" -s 2 -c 2 -f -o $O/r02c_fan21_scatter $K > $O/r02c_ncu_part.log 2>&1; echo "ncu scatter rc=$?" >> $O/r02c_status.txt
unset DNAGPU_LIB
# the bucket count at its natural bucket size (100 Mbp, k = 21: 2^16 buckets of ~ 1526 keys)
K="python tools/ksweep.py --n-bases 100000000 --seed 2 --ks 21 --reps 1"
$K > $O/r02c_ksweep_c2.json 2>&1 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_count_buckets_bins" -s 1 -c 1 \
    -f -o $O/r02c_c2_count_bins $K > $O/r02c_ncu_count.log 2>&1; echo "ncu count rc=$?" >> $O/r02c_status.txt
F="python bench.py --workload c3 --n-bases 1500000000 --steps 1 --warmup 3 --e2e-steps 1 --cpu-sample 1000000"
$F > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"k_filter_collect" -s 3 -c 1 -f -o $O/r02c_c3_filter_sa $F > $O/r02c_ncu_f1.log 2>&1
$F --planes > /dev/null 2>&1 && ncu --set full --import-source on --clock-control none -k regex:"k_filter_collect" -s 3 -c 1 -f -o $O/r02c_c3_filter_planes $F --planes > $O/r02c_ncu_f2.log 2>&1
echo "ncu filter rc=$?" >> $O/r02c_status.txt
cat $O/r02c_status.txt
python - <<'PY'
import json
for f in ("c3","c3_planes","c2","c5_sweep","c1","reference","reference_c1"):
    try:
        d=json.load(open(f"gpurun_out/r02c_bench_{f}.json"))
        print(f, round(d["value"],4), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],4), (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "failed", e)
PY
