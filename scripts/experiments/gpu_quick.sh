#!/bin/bash
# quick loop: mlp/module tests + bench (planned and unplanned)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mlp.py tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -6
for plan in 0; do
  echo "== PRB_ENABLE_PLAN=$plan"
  PRB_ENABLE_PLAN=$plan timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/bench_plan$plan.json > gpurun_out/bench_plan$plan.log 2>&1 || tail -5 gpurun_out/bench_plan$plan.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_plan$plan.json'))
print(round(d['value'],1),'scenes/s', round(d['ms_per_step'],3),'ms  e2e',round(d['e2e']['value'],1),'single',round(d['single_batch']['ms_per_step'],3), {k['name'][:24]:round(k['ms_per_step'],3) for k in d['kernels']})
PY
done
