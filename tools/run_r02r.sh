#!/bin/bash
# GPU run r02r (1 GPU, final build): whole -m gpu suite, smoke, the bench lines of every workload, the reference arm on a
# 100 Mbp sample, ncu: launch list of the headline command, instruction counts + --set full of the multi-GPU front end.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02r}
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${TAG}_status.txt
timeout 900 python bench.py 2> $O/${TAG}_bench_c4.err | grep '^{' > $O/${TAG}_bench_c4.json; echo "bench c4 rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
for w in c2 c3 c1; do
  timeout 600 python bench.py --workload $w --cpu-sample 1000000 2> $O/${TAG}_bench_$w.err | grep '^{' > $O/${TAG}_bench_$w.json; echo "bench $w rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
done
timeout 900 python bench.py --impl reference --cpu-large-sample 100000000 --steps 1 --warmup 0 2> $O/${TAG}_bench_reference_100mbp.err | grep '^{' > $O/${TAG}_bench_reference_100mbp.json; echo "reference rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 80 --csv --log-file $O/${TAG}_c4_launches_dram.csv \
    python bench.py --steps 1 --warmup 3 --e2e-steps 1 --no-extract --cpu-sample 100000 > $O/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?" >> $O/${TAG}_status.txt
for g in 8 4 2 3; do
  python tools/owned_once.py --parts $g > $O/${TAG}_owned_once_g$g.json 2> $O/${TAG}_owned_once_g$g.err
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
      --clock-control none -k regex:k_collect_owned -s 1 -c 1 --csv --log-file $O/${TAG}_collect_owned_g${g}_inst.csv python tools/owned_once.py --parts $g > /dev/null 2>&1; echo "ncu inst g=$g rc=$?" >> $O/${TAG}_status.txt
done
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_collect_owned -s 1 -c 1 -f -o $O/${TAG}_collect_owned_g8_full python tools/owned_once.py --parts 8 > $O/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; tail -3 $O/${TAG}_pytest_gpu.log; cat $O/${TAG}_owned_once_g*.json
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "value", d.get("value"), d.get("unit"), "ms", d.get("ms_per_step"), "e2e", (d.get("e2e") or {}).get("value"), "roofline", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
grep -h "k_collect_owned" $O/${TAG}_collect_owned_g*_inst.csv | cut -c1-400
