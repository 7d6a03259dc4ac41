#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_tests3a.log
tail -6 gpurun_out/r2_tests3a.log
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_mlp.py 2>&1 | tail -30 > gpurun_out/r2_tests3b.log
tail -12 gpurun_out/r2_tests3b.log
for pool in 0 1; do
  PRB_MLP_POOL=$pool PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline --min-seconds 0.3 --profile-out gpurun_out/r2_bench_lean$pool.json > gpurun_out/r2_bench_lean$pool.log 2>&1 || tail -5 gpurun_out/r2_bench_lean$pool.log
done
python - <<'PY'
import json
for pl in (0, 1):
    try:
        d = json.load(open("gpurun_out/r2_bench_lean%d.json" % pl))
        print("pool", pl, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"], 3))
        for k in d["kernels"]:
            print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
    except Exception as e:
        print("pool", pl, "failed", e)
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:mlp_pipe_kernel -s 32 -c 16 --csv --log-file gpurun_out/r2_lean_launches.csv python scripts/one_forward.py 3 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:roipool -c 12 --csv --log-file gpurun_out/r2_roipool_launches.csv python scripts/bench_ops.py > /dev/null 2>&1
timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
cat gpurun_out/r2_rcnn_stage.json | head -c 1500
