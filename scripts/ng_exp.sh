#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "fps" -x 2>&1 | tail -3
for v in base ng2occ2; do
  echo "== $v"
  if [ $v = ng2occ2 ]; then export PRB_MLP_NG=2 PRB_MLP_NG2_OCC2=1; fi
  PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_levels_$v.json > gpurun_out/bench_levels_$v.log 2>&1 || tail -5 gpurun_out/bench_levels_$v.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_levels_$v.json'))
print(round(d['value']), 'single', round(d['single_batch']['ms_per_step'],3))
for k in d['kernels']: print('%-60s %.4f' % (k['name'][:60], k['ms_per_step']))
PY
done
