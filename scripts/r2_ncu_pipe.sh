#!/bin/bash
# ncu --set full of the chain launches of ONE forward (after 2 warm-up forwards = 32 chain launches skipped)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -s 32 -c 16 -o gpurun_out/r2_pipe_ncu -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_pipe.log 2>&1
tail -3 gpurun_out/r2_ncu_pipe.log
ls -la gpurun_out/r2_pipe_ncu.ncu-rep
ncu -i gpurun_out/r2_pipe_ncu.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu_raw.csv 2>/dev/null
