#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests7.log
tail -8 gpurun_out/r2_tests7.log | cut -c1-200
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["launch"][:44].ljust(44), d["ms"], "IA", d["issuerA"], "IB", d["issuerB"], "G", d["gather0"], "E", d["epi0"], "P1", d["prod1"])
'
timeout 200 python scripts/pipe_trace.py 2>&1 | grep "^{" | tee gpurun_out/r2_pipe_trace7.jsonl | python -c "$show" | cut -c1-520
echo "--- zs=64 nbuf=1"
timeout 200 python scripts/pipe_trace.py mlp_zs=64 mlp_nbuf=1 2>&1 | tail -30 | grep "^{\|rror" | python -c "$show" | cut -c1-520
for pool in 0 2; do
  PRB_MLP_POOL=$pool timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage_pool$pool.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
  echo "rcnn pool=$pool"; head -c 900 gpurun_out/r2_rcnn_stage_pool$pool.json; echo
done
PRB_PROF_DETAIL=1 timeout 900 python bench.py --profile-out gpurun_out/r2_bench_b7.json > gpurun_out/r2_bench_b7.log 2>&1 || tail -5 gpurun_out/r2_bench_b7.log
tail -1 gpurun_out/r2_bench_b7.log | cut -c1-3000
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b7.json"))
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
except Exception as e:
    print("bench failed", e)
PY
