#!/bin/bash
# round 2, first GPU pass: tests on the options/pipeline rework, bench with and without CUDA graphs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | tail -15 > gpurun_out/r2_tests1.log
tail -5 gpurun_out/r2_tests1.log
timeout 400 python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_bench_g1.json > gpurun_out/r2_bench_g1.log 2>&1 || tail -20 gpurun_out/r2_bench_g1.log
timeout 300 python bench.py --steps 20 --warmup 5 --graphs 0 --no-ref-cuda --no-cpu-baseline --profile-out gpurun_out/r2_bench_g0.json > gpurun_out/r2_bench_g0.log 2>&1 || tail -20 gpurun_out/r2_bench_g0.log
python - <<'PY'
import json
for g in (1, 0):
    try:
        d = json.load(open("gpurun_out/r2_bench_g%d.json" % g))
        print("graphs", g, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "feat", round(d["e2e_features"]["value"]),
              "single", round(d["single_batch"]["ms_per_step"], 3), "rep", d["repeats"]["n"], d["vs_ref_cuda"], d["ref_cuda"] and {k: v for k, v in d["ref_cuda"].items() if k in ("single_stream", "pipelined", "unavailable")})
    except Exception as e:
        print("graphs", g, "failed", e)
PY
