#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -8 | cut -c1-250
PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b20.json > gpurun_out/r2_bench_b20.log 2>&1 || tail -5 gpurun_out/r2_bench_b20.log
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_b20.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"].get("ms_per_step_planned"))
for k in d["kernels"]:
    print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
for c in d["chain_plans"]:
    if c["in"]=="sa" and c["out"]=="sa_max": print("   plan", c["nsample"], c["np"], c["build"])
PY
