#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests12.log
tail -12 gpurun_out/r2_tests12.log | cut -c1-250
for pool in 0 2; do
PRB_MLP_POOL=$pool PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b12_$pool.json > gpurun_out/r2_bench_b12.log 2>&1 || tail -5 gpurun_out/r2_bench_b12.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b12_$pool.json"))
    print("pool $pool value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"].get("ms_per_step_planned"))
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
except Exception as e:
    print("bench failed", e)
PY
done
for pool in 0 2; do
PRB_MLP_POOL=$pool timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage12_$pool.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
echo "rcnn pool $pool"; head -c 900 gpurun_out/r2_rcnn_stage12_$pool.json; echo
done
