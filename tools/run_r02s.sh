#!/bin/bash
# GPU run r02s (2 GPUs): NVLink byte counters of the driver (nvidia-smi nvlink -gt d: cumulative data Tx / Rx per link)
# read before and after STEPS owner-restricted counts of the whole 3.1 Gbp sequence on a 2-GPU context, so that the bytes
# that crossed NVLink per step can be set against the 0.25 B/base * (G-1)/G the design claims.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02s}; STEPS=20
snap() { for i in 0 1; do echo "== GPU $i"; nvidia-smi nvlink -gt d -i $i; done; }
nvidia-smi nvlink -s -i 0 > $O/${TAG}_nvlink_status.txt 2>&1
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1
snap > $O/${TAG}_nvlink_before.txt 2>&1
./dna-sequences-pg-extension_b200/dnagpu_bench --bases 3100000000 --k 31 --seed 4 --steps $STEPS --host --gpus 2 > $O/${TAG}_cbench_n2.json 2> $O/${TAG}_cbench.err; echo "cbench rc=$?" > $O/${TAG}_status.txt
snap > $O/${TAG}_nvlink_after.txt 2>&1
python - <<PY
import re, json
def read(f):
    out, gpu = {}, None
    for l in open(f):
        m = re.match(r"== GPU (\d+)", l)
        if m: gpu = int(m.group(1)); out[gpu] = {"tx": 0, "rx": 0, "links": 0}; continue
        m = re.search(r"Link (\d+): Data Tx: (\d+) KiB", l)
        if m and gpu is not None: out[gpu]["tx"] += int(m.group(2)) * 1024; out[gpu]["links"] += 1
        m = re.search(r"Link (\d+): Data Rx: (\d+) KiB", l)
        if m and gpu is not None: out[gpu]["rx"] += int(m.group(2)) * 1024
    return out
b, a = read("gpurun_out/${TAG}_nvlink_before.txt"), read("gpurun_out/${TAG}_nvlink_after.txt")
counts = $STEPS + 1  # + the warm-up call
n_bases, G = 3100000000, 2
res = {"counts": counts, "expected_rx_bytes_per_gpu_per_count": 0.25 * n_bases * (G - 1) / G, "gpus": {}}
for g in a:
    if g in b:
        res["gpus"][g] = {"links": a[g]["links"], "rx_bytes_per_count": (a[g]["rx"] - b[g]["rx"]) / counts, "tx_bytes_per_count": (a[g]["tx"] - b[g]["tx"]) / counts}
print(json.dumps(res))
open("gpurun_out/${TAG}_nvlink_bytes.json", "w").write(json.dumps(res, indent=1))
PY
cat $O/${TAG}_status.txt; cat $O/${TAG}_cbench_n2.json; head -8 $O/${TAG}_nvlink_after.txt
