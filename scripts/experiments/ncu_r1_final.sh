#!/bin/bash
# launch list of one sequential bench pass + full captures of the pruned FPS, the 3-NN grid walk and one chain launch
mkdir -p gpurun_out
B="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_pruned -s 3 -c 1 -o gpurun_out/prof_fps_pruned $B > gpurun_out/ncu_full1.log 2>&1; echo "fps rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:three_nn_grid -s 12 -c 1 -o gpurun_out/prof_three_nn $B > gpurun_out/ncu_full2.log 2>&1; echo "3nn rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 63 -c 2 -o gpurun_out/prof_mlp_final $B > gpurun_out/ncu_full3.log 2>&1; echo "mlp rc=$?"
