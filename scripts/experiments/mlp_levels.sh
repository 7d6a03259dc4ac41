#!/bin/bash
# per-level device times of the fused chain kernels (sequential pass, CUDA events)
mkdir -p gpurun_out
PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/bench_levels.json > gpurun_out/bench_levels.log 2>&1 || tail -5 gpurun_out/bench_levels.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_levels.json'))
for k in d['kernels']: print('%-70s %.4f' % (k['name'][:70], k['ms_per_step']))
PY
