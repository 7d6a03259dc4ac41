#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -s 32 -c 16 -o gpurun_out/r2_pipe_ncu -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_pipe.log 2>&1
tail -2 gpurun_out/r2_ncu_pipe.log
ncu -i gpurun_out/r2_pipe_ncu.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu_raw.csv 2>/dev/null
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r2_tests2.log
tail -25 gpurun_out/r2_tests2.log
timeout 300 python scripts/bench_ops.py > gpurun_out/r2_bench_ops.json 2> gpurun_out/r2_bench_ops.err || tail -5 gpurun_out/r2_bench_ops.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ops.json')); print(json.dumps(d['roipool3d_C4'], indent=1))"
