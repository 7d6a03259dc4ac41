#!/bin/bash
mkdir -p gpurun_out
for ns in 400 1000 2000 4000; do
PRB_MLP_LAZY_NS=$ns PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b22_$ns.json > gpurun_out/r2_bench_b22.log 2>&1 || tail -5 gpurun_out/r2_bench_b22.log
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_b22_$ns.json"))
k={x["name"]:x["ms_per_step"] for x in d["kernels"]}
print("lazy_ns $ns value", round(d["value"]), "SA", round(d["kernels"][0]["ms_per_step"],4), "FP", round(d["kernels"][1]["ms_per_step"],4), {n[7:22]: round(v,4) for n,v in k.items() if n.startswith("sa_mlp 65536") or n.startswith("sa_mlp 16384")})
PY
done
