#!/bin/bash
# three_nn tests, then the cell-edge factor sweep (family time of the sequential pass)
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --no-header -p no:cacheprovider -k "three_nn or fp_module" -x 2>&1 | tail -4
for f in 1.0 1.3 1.6 2.0; do
  echo "== PRB_NN_CELL=$f"
  PRB_NN_CELL=$f PRB_GRID_DEBUG=1 timeout 100 python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline 2>&1 | grep "three_nn_grid" | tail -2
  PRB_NN_CELL=$f timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(round(d['value']), 'e2e', round(d['e2e']['value']), 'single', round(d['single_batch']['ms_per_step'],3), {k['name'][:10]:round(k['ms_per_step'],3) for k in d['kernels']})"
done
