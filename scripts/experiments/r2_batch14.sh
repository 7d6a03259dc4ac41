#!/bin/bash
mkdir -p gpurun_out
for fill in 1 0; do
PRB_MLP_FILL=$fill PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b14_$fill.json > gpurun_out/r2_bench_b14.log 2>&1 || tail -5 gpurun_out/r2_bench_b14.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b14_$fill.json"))
    print("fill $fill value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"].get("ms_per_step_planned"))
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
except Exception as e:
    print("bench failed", e)
PY
done
