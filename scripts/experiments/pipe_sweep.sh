#!/bin/bash
for rep in 1 2; do for cs in 0 2; do for f in 3 4 6; do
  PRB_FPS_CS=$cs timeout 120 python bench.py --inflight $f --steps 80 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('cs',$cs,'inflight',$f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3))"
done; done; done
