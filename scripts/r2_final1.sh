#!/bin/bash
# final single-GPU records: default bench line, reference arm, ncu launch list of the same bench command
mkdir -p gpurun_out
timeout 1200 python bench.py --profile-out gpurun_out/r2_bench_final.json > gpurun_out/r2_bench_final.log 2>&1 || tail -5 gpurun_out/r2_bench_final.log
tail -1 gpurun_out/r2_bench_final.log | cut -c1-1500
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_arm.log 2>&1
tail -1 gpurun_out/r2_bench_reference_arm.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train --no-rcnn --min-seconds 0.01 > gpurun_out/r2_ncu_list.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/r2_launches_final.csv
