#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_mlp.py -m gpu -q --no-header -p no:cacheprovider 2>&1 | tail -3
for g in 0 1 2; do
  PRB_MLP_GATHER=$g timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/bench_g$g.json > gpurun_out/bench_g$g.log 2>&1 || tail -3 gpurun_out/bench_g$g.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_g$g.json'))
print('gather',$g, round(d['value'],1),'scenes/s', round(d['ms_per_step'],3),'ms single',round(d['single_batch']['ms_per_step'],3), {k['name'][:22]:round(k['ms_per_step'],3) for k in d['kernels']})
PY
done
