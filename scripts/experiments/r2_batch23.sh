#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -6 | cut -c1-250
for r in 1 0; do
PRB_MLP_RESIDENT=$r PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b23_$r.json > gpurun_out/r2_bench_b23.log 2>&1 || tail -5 gpurun_out/r2_bench_b23.log
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_b23_$r.json"))
k={x["name"]:x["ms_per_step"] for x in d["kernels"]}
print("resident $r value", round(d["value"]), "SA", round(d["kernels"][0]["ms_per_step"],4), "FP", round(d["kernels"][1]["ms_per_step"],4), {n[7:22]: round(v,4) for n,v in k.items() if n.startswith("sa_mlp") or n.startswith("fp_mlp")})
PY
done
