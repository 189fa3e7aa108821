#!/bin/bash
# GPU run r02u (1 GPU): ownership tests with the widened parameter lists (G = 4, tiny inputs at G = 8, both fallbacks)
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02u}
timeout 900 python -m pytest tests/test_gpu_owned.py tests/test_gpu_lifetime.py -m gpu -q -p no:cacheprovider > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" > $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; tail -3 $O/${TAG}_pytest.log
