#!/bin/bash
# GPU run r02t (2 GPUs): the two commands the driver runs for a scaling point, verbatim (no extra flags), reference arm first.
cd "$(dirname "$0")/.."
O=gpurun_out; TAG=${1:-r02t}
T0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
   > $O/${TAG}_reference_n2.out 2> $O/${TAG}_reference_n2.err; echo "reference rc=$? $(( $(date +%s) - T0 )) s" > $O/${TAG}_status.txt
T0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 25 --warmup 5 \
   > $O/${TAG}_bench_n2.out 2> $O/${TAG}_bench_n2.err; echo "bench rc=$? $(( $(date +%s) - T0 )) s" >> $O/${TAG}_status.txt
cat $O/${TAG}_status.txt; grep -c '^{' $O/${TAG}_reference_n2.out $O/${TAG}_bench_n2.out; grep '^{' $O/${TAG}_reference_n2.out | cut -c1-600; grep '^{' $O/${TAG}_bench_n2.out | cut -c1-900; tail -3 $O/${TAG}_bench_n2.err
