#!/bin/bash
# GPU run r02b (2 GPUs): multi-GPU tests, bench N=2 in the gather and peer forms, the C harness on 2 GPUs
cd "$(dirname "$0")/.."
O=gpurun_out
nvidia-smi -L > $O/r02b_gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "multi or ring or harness or two_gpu or peer" -p no:cacheprovider > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> $O/r02b_pytest.log
for ex in gather peer; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 3 --cpu-sample 1000000 --no-extract --exchange $ex \
    > $O/r02b_bench_c4_n2_$ex.json 2> $O/r02b_bench_c4_n2_$ex.err; echo "bench $ex rc=$?" >> $O/r02b_pytest.log
done
timeout 300 ./dna-sequences-pg-extension_b200/dnagpu_bench --bases 3100000000 --k 31 --seed 4 --steps 5 --host --gpus 2 > $O/r02b_cbench_n2.json 2> $O/r02b_cbench_n2.err; echo "cbench rc=$?" >> $O/r02b_pytest.log
timeout 300 ./dna-sequences-pg-extension_b200/dnagpu_bench --bases 3100000000 --k 31 --seed 4 --steps 5 --host --gpus 1 > $O/r02b_cbench_n1.json 2>> $O/r02b_cbench_n2.err
tail -4 $O/r02b_pytest.log
python - <<'PY'
import json
for f in ("r02b_bench_c4_n2_gather","r02b_bench_c4_n2_peer","r02b_cbench_n2","r02b_cbench_n1"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, round(d.get("value", d.get("gkmer_s")),2), round(d["ms_per_step"],2), d.get("e2e",{}).get("value"), {k:round(v["ms"]/v["launches"],2) for k,v in d.get("kernels",{}).items()} if isinstance(d.get("kernels"),dict) else "")
    except Exception as e:
        print(f, "failed", e)
PY
