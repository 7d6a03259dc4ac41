#!/bin/bash
# what the driver runs at round end: GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --profile-out gpurun_out/bench_final.json > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_refarm.log 2>&1; tail -1 gpurun_out/bench_refarm.log | cut -c1-200
