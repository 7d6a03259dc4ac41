#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"three_nn_grid_kernel|ball_query_grid_kernel" -s 8 -c 4 -o gpurun_out/r2_nn_ncu -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_nn.log 2>&1
tail -2 gpurun_out/r2_ncu_nn.log
ncu -i gpurun_out/r2_nn_ncu.ncu-rep --page details --csv > gpurun_out/r2_nn_ncu_details.csv 2>/dev/null
ncu -i gpurun_out/r2_nn_ncu.ncu-rep --page raw --csv > gpurun_out/r2_nn_ncu_raw.csv 2>/dev/null
ls -la gpurun_out/r2_nn_ncu*
