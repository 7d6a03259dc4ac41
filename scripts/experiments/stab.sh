#!/bin/bash
# full GPU suite, then the default bench three times (stability of value vs e2e)
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -4
for r in 1 2 3; do
  timeout 200 python bench.py --no-cpu-baseline --profile-out gpurun_out/bench_stab$r.json > gpurun_out/bench_stab$r.log 2>&1 || tail -5 gpurun_out/bench_stab$r.log
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_stab$r.json'))
print(round(d['value'],1),'scenes/s', round(d['ms_per_step'],3),'ms  e2e',round(d['e2e']['value'],1),'single',round(d['single_batch']['ms_per_step'],3), {k['name'][:24]:round(k['ms_per_step'],3) for k in d['kernels']}, d['gpu_launches'])
PY
done
