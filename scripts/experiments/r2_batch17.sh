#!/bin/bash
mkdir -p gpurun_out
for sms in 0 132 124 140; do
PRB_MLP_SMS=$sms timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b17_$sms.json > gpurun_out/r2_bench_b17.log 2>&1 || tail -5 gpurun_out/r2_bench_b17.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b17_$sms.json"))
    print("mlp_sms $sms value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"],3), "SA", round(d["kernels"][0]["ms_per_step"],4), "FP", round(d["kernels"][1]["ms_per_step"],4))
except Exception as e:
    print("bench failed", e)
PY
done
