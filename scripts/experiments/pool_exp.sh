#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_mlp.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 100 python scripts/mlp_trace.py 2>&1 | grep "cycles/phase" | cut -c1-330
PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/bench_pool.json > gpurun_out/bench_pool.log 2>&1 || tail -5 gpurun_out/bench_pool.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_pool.json'))
print(round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3), ' '.join('%s=%.3f' % (k['name'].split(' ')[0][:3]+k['name'].split(' ')[1][:12] if k['name'].startswith(('sa_mlp ','fp_mlp ')) else k['name'][:12], k['ms_per_step']) for k in d['kernels']))
PY
