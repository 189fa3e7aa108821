#!/bin/bash
# GPU run r02o (2 GPUs): ownership test as compare + select (main build) against IMAD.HI + funnel shift (tuning build,
# -DDNAGPU_OWN_CARRY): owned tests on both, the N=8/4/2 traffic probe on both.  The variant measured slower
# (profiles/README.md, r02o) and was removed from the source afterwards; this script records how it was run.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02o}
T=$PWD/dna-sequences-pg-extension_b200/libdnagpu_tuning.so
timeout 900 python -m pytest tests/test_gpu_owned.py tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest_main.log 2>&1; echo "pytest main rc=$?" > $O/${TAG}_status.txt
DNAGPU_LIB=$T timeout 900 python -m pytest tests/test_gpu_owned.py tests/test_gpu_multi.py tests/test_abi.py -q -p no:cacheprovider > $O/${TAG}_pytest_carry.log 2>&1; echo "pytest carry rc=$?" >> $O/${TAG}_status.txt
for parts in 8 4 2; do
  for v in main carry; do
    L=""; [ $v = carry ] && L=$T
    DNAGPU_LIB=$L timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$parts \
      tools/probe_collect.py --parts $parts 2> $O/${TAG}_probe_p${parts}_$v.err | grep '^{' > $O/${TAG}_probe_p${parts}_$v.jsonl; echo "probe $parts $v rc=${PIPESTATUS[0]}" >> $O/${TAG}_status.txt
  done
done
cat $O/${TAG}_status.txt; tail -2 $O/${TAG}_pytest_main.log $O/${TAG}_pytest_carry.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_probe_*.jsonl")):
    for l in open(f):
        d=json.loads(l); print(f.split("/")[-1], "rank", d["rank"], "peer", d["peer_fraction"], "collect", d["kernels_ms"]["collect_owned"], "total", d["owned_total"], d["distinct"], d["unique"])
PY
