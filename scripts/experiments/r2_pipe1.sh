#!/bin/bash
# pipelined chain kernel: correctness first (short timeouts: a deadlock must not hold the box), then per-level times
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_mlp.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/r2_pipe_tests.log
rc=${PIPESTATUS[0]}
tail -12 gpurun_out/r2_pipe_tests.log
for pl in 1 0; do
  PRB_MLP_PIPELINE=$pl PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline --min-seconds 0.3 --profile-out gpurun_out/r2_bench_pipe$pl.json > gpurun_out/r2_bench_pipe$pl.log 2>&1 || tail -5 gpurun_out/r2_bench_pipe$pl.log
done
python - <<'PY'
import json
for pl in (1, 0):
    try:
        d = json.load(open("gpurun_out/r2_bench_pipe%d.json" % pl))
        print("pipeline", pl, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"], 3))
        for k in d["kernels"]:
            print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
    except Exception as e:
        print("pipeline", pl, "failed", e)
PY
