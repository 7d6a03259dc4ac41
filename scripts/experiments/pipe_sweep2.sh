#!/bin/bash
# pipelined throughput: MLP grid size (SMs left to the chain kernels) x batches in flight, pruned FPS on
for sms in 148 132 116 100; do for f in 4 6; do
  PRB_MLP_SMS=$sms timeout 120 python bench.py --inflight $f --steps 60 --warmup 8 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print('mlp_sms',$sms,'inflight',$f, round(d['value']), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3))"
done; done
