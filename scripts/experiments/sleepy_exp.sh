#!/bin/bash
mkdir -p gpurun_out
for sl in 3 2 0; do
  echo "== PRB_MLP_SLEEPY=$sl"
  PRB_MLP_SLEEPY=$sl timeout 100 python scripts/mlp_trace.py 2>&1 | grep "cycles/phase" | cut -c1-330
  PRB_MLP_SLEEPY=$sl PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_sl$sl.json > gpurun_out/bench_sl$sl.log 2>&1 || tail -5 gpurun_out/bench_sl$sl.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_sl$sl.json'))
print(round(d['value']), 'single', round(d['single_batch']['ms_per_step'],3), ' '.join('%s=%.3f' % (k['name'].split(' ')[0][:3]+k['name'].split(' ')[1][:12] if k['name'].startswith(('sa_mlp ','fp_mlp ')) else k['name'][:12], k['ms_per_step']) for k in d['kernels']))
PY
done
