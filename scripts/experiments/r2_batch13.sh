#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests13.log
tail -12 gpurun_out/r2_tests13.log | cut -c1-250
for gl in 0 1; do
PRB_GRID_LISTS=$gl PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b13_$gl.json > gpurun_out/r2_bench_b13.log 2>&1 || tail -5 gpurun_out/r2_bench_b13.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b13_$gl.json"))
    print("grid_lists $gl value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"])
    for k in d["kernels"][:6]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
except Exception as e:
    print("bench failed", e)
PY
done
