#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --profile-out gpurun_out/r2_bench_n2.json > gpurun_out/r2_bench_n2.log 2>&1
echo "rc=$?"; tail -1 gpurun_out/r2_bench_n2.log | cut -c1-800
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_n2.json"))
    print("N=2 value", d["value"], "e2e", d["e2e"]["value"])
    print("strong", json.dumps(d["strong_scaling"])[:600])
    print("train", json.dumps({k:v for k,v in d["train_step"].items() if k!="what"})[:600])
except Exception as e:
    print("failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_n2.log 2>&1
echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref_n2.log | cut -c1-300
timeout 300 python -m pytest tests -m gpu -q -p no:cacheprovider -k "train or nccl or reducer" 2>&1 | tail -3
