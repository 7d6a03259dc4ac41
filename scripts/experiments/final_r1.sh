#!/bin/bash
# round-1 final: full GPU suite, default bench (with the CPU baseline leg), reference arm, launch list + full captures
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x 2>&1 | tail -3
timeout 400 python bench.py --profile-out gpurun_out/bench_final.json > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_refarm.log 2>&1; tail -1 gpurun_out/bench_refarm.log | cut -c1-400
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
B="python bench.py --inflight 1 --steps 1 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv $B > gpurun_out/ncu_bench.log 2>&1; echo "list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mlp_chain -s 63 -c 2 -o gpurun_out/prof_mlp_final $B > gpurun_out/ncu_full3.log 2>&1; echo "mlp rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fps_pruned -s 3 -c 1 -o gpurun_out/prof_fps_pruned $B > gpurun_out/ncu_full1.log 2>&1; echo "fps rc=$?"
