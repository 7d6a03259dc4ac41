#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --profile-out gpurun_out/r2_bench_n2b.json > gpurun_out/r2_bench_n2b.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_bench_n2b.log | cut -c1-400
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/r2_bench_n2b.json"))
    print("N=2 value", d["value"], "e2e", d["e2e"]["value"])
    print("strong", json.dumps(d["strong_scaling"])[:500])
    print("train", json.dumps({k:v for k,v in d["train_step"].items() if k!="what"})[:900])
    print("eval", json.dumps({k:v for k,v in d["eval_e2e"].items() if k!="what"})[:600])
except Exception as e:
    print("failed", e)
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r2_bench_ref_n2b.log 2>&1
echo "ref rc=$?"; tail -1 gpurun_out/r2_bench_ref_n2b.log | cut -c1-200
