#!/bin/bash
# One gpurun call: parity tests, golden vectors from the reference kernels, bench, launch list.
# Every step has its own timeout so a hung kernel cannot eat the whole lease.  Logs -> gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "== ops tests";   timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "ops rc=$?"
tail -5 gpurun_out/pytest_ops.log
echo "== golden";      timeout 300 python oracle/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"
echo "== mlp tests";   timeout 900 python -m pytest tests/test_gpu_mlp.py -m gpu -q -rf --no-header -p no:cacheprovider > gpurun_out/pytest_mlp.log 2>&1; echo "mlp rc=$?"
tail -5 gpurun_out/pytest_mlp.log
echo "== smoke";       timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== ref cuda"; timeout 600 python oracle/bench_ref_cuda.py > gpurun_out/bench_ref_cuda.log 2>&1; echo "refcuda rc=$?"; tail -1 gpurun_out/bench_ref_cuda.log | cut -c1-800
echo "== bench";       timeout 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/bench_profile.json > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | cut -c1-1500
if [ "$1" == "ncu" ]; then
  echo "== ncu launch list"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
fi
echo done
