#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_mlp.py tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -12 | cut -c1-250
for w in 0 1; do
PRB_NN_WALK=$w timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b16_$w.json > gpurun_out/r2_bench_b16.log 2>&1 || tail -5 gpurun_out/r2_bench_b16.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b16_$w.json"))
    print("nn_walk $w value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"].get("ms_per_step_planned"))
    for k in d["kernels"][:6]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
except Exception as e:
    print("bench failed", e)
PY
done
