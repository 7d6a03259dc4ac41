#!/bin/bash
# pruned FPS also for the 4096-point level (PRB_FPS_PRUNE=2), and the pipeline depth
for pr in 1 2; do for f in 6 8; do
  PRB_FPS_PRUNE=$pr timeout 150 python bench.py --inflight $f --steps 40 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('prune',$pr,'inflight',$f, round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3), {k['name'][:10]:round(k['ms_per_step'],3) for k in d['kernels']})"
done; done
