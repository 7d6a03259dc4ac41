#!/bin/bash
mkdir -p gpurun_out
PRB_MLP_ATMEM=1 timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for v in 0 1; do
  PRB_MLP_ATMEM=$v PRB_PROF_DETAIL=1 timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/bench_atmem$v.json > gpurun_out/bench_atmem$v.log 2>&1 || tail -3 gpurun_out/bench_atmem$v.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_atmem$v.json'))
print('atmem',$v, round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3), ' '.join('%s=%.3f' % (k['name'].split(' ')[0][:3]+k['name'].split(' ')[1][:12] if k['name'].startswith(('sa_mlp ','fp_mlp ')) else k['name'][:12], k['ms_per_step']) for k in d['kernels']))
PY
done
