#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b21.json > gpurun_out/r2_bench_b21.log 2>&1 || tail -5 gpurun_out/r2_bench_b21.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_b21.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]))
for c in d["chain_plans"]: print("   ", c["in"], c["out"], c["nsample"], c["np"], "->", c["build"][:12], c["measured_us"])
PY
