#!/bin/bash
# bench the backbone with different row-group counts of the MLP chain kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for ng in 0 1 2 3; do
  echo "== PRB_MLP_NG=$ng"
  PRB_MLP_NG=$ng timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_ng$ng.json > gpurun_out/bench_ng$ng.log 2>&1
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_ng$ng.json'))
print(round(d['value'],1),'scenes/s', round(d['ms_per_step'],3),'ms', {k['name'][:22]:round(k['ms_per_step'],3) for k in d['kernels']})
PY
done
