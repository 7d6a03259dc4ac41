#!/bin/bash
# full ncu capture of the tensor-core chain (SA2 scale 1 + SA3 scale 0 launches) and of the FPS cluster kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 54 -c 2 -o gpurun_out/prof_mlp python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "mlp rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fps_rank -s 12 -c 1 -o gpurun_out/prof_fps python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1; echo "fps rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "list rc=$?"
python bench.py --steps 30 --warmup 5 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | cut -c1-400
