#!/bin/bash
# final-state ncu records: (1) launch list of the bench command, (2) --set full of every kernel family once (scripts/ops_once.py)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches_s2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-ref-cuda --no-train --min-seconds 0.01 > gpurun_out/r2_ncu_list_s2.log 2>&1
echo "list rc=$?"; wc -l gpurun_out/r2_launches_s2.csv
timeout 900 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/r2_ops_ncu -f python scripts/ops_once.py > gpurun_out/r2_ncu_ops.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu_ops.log
ncu -i gpurun_out/r2_ops_ncu.ncu-rep --page raw --csv > gpurun_out/r2_ops_ncu_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_ops_ncu_raw.csv gpurun_out/r2_ncu_ops_summary.csv
ls -la gpurun_out/r2_ops_ncu.ncu-rep
# the merge back is limited to 64 MiB: keep the report only if it is small
find gpurun_out -name r2_ops_ncu.ncu-rep -size +30M -delete
