#!/bin/bash
# launch list of one bench pass (serialised, cold cache: compare shares) + optional full capture of launches matching $1 (skip $2, count $3)
mkdir -p gpurun_out
PRB_DISABLE_PLAN=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "list rc=$?"
if [ -n "$1" ]; then
  PRB_DISABLE_PLAN=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-0} -c ${3:-1} -o gpurun_out/prof_$1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
  echo "full rc=$?"
fi
