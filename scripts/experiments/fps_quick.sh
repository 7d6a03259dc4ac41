#!/bin/bash
# FPS parity tests, then the bench with and without the pruned kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k fps -x 2>&1 | tail -8
for pr in 0 1; do
  echo "== PRB_FPS_PRUNE=$pr"
  PRB_FPS_PRUNE=$pr timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/bench_prune$pr.json > gpurun_out/bench_prune$pr.log 2>&1 || tail -5 gpurun_out/bench_prune$pr.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_prune$pr.json'))
print(round(d['value'],1),'scenes/s', round(d['ms_per_step'],3),'ms  e2e',round(d['e2e']['value'],1),'single',round(d['single_batch']['ms_per_step'],3), {k['name'][:24]:round(k['ms_per_step'],3) for k in d['kernels']})
PY
done
