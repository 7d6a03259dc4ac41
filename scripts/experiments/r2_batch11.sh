#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests11.log
tail -12 gpurun_out/r2_tests11.log | cut -c1-250
timeout 300 python scripts/bench_ops.py 2>/dev/null > gpurun_out/r2_bench_ops11.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ops11.json'))
for k,v in d.items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','ms_reference','speedup','ms_with_canonical','frac','ms_reference_kernels','ms_exhaustive')})"
timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage11.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
head -c 700 gpurun_out/r2_rcnn_stage11.json; echo
for mr in 131072 65536; do
PRB_FP_PROJECT_MIN_ROWS=$mr PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b11_$mr.json > gpurun_out/r2_bench_b11.log 2>&1 || tail -5 gpurun_out/r2_bench_b11.log
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b11_$mr.json"))
    print("min_rows $mr value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"].get("ms_per_step_planned"))
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
except Exception as e:
    print("bench failed", e)
PY
done
