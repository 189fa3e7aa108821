#!/bin/bash
# GPU run r02y (2 GPUs): the PG glue with a backend context over all GPUs (DNAGPU_DEVICES)
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02y}
timeout 500 python -m pytest tests/test_gpu_pg_glue.py -m gpu -q -p no:cacheprovider -k "all_gpus or pushdown_functions" > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; tail -15 $O/${TAG}_pytest.log
