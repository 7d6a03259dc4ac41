#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests6.log
tail -15 gpurun_out/r2_tests6.log | cut -c1-200
PRB_PROF_DETAIL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --min-seconds 0.3 --profile-out gpurun_out/r2_bench_b6.json > gpurun_out/r2_bench_b6.log 2>&1 || tail -5 gpurun_out/r2_bench_b6.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b6.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"], 3))
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
except Exception as e:
    print("bench failed", e)
PY
timeout 200 python scripts/pipe_trace.py 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['launch'][:44].ljust(44), d['ms'], 'IA', d['issuerA'], 'IB', d['issuerB'], 'G', d['gather0'], 'E', d['epi0'], 'P1', d['prod1'])
" | cut -c1-420
echo "--- zs=128 nbuf=1"
timeout 200 python scripts/pipe_trace.py mlp_zs=128 mlp_nbuf=1 2>&1 | grep "^{" | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['launch'][:44].ljust(44), d['ms'])
"
timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
cat gpurun_out/r2_rcnn_stage.json | head -c 1200; echo
timeout 300 python scripts/bench_ops.py 2>/dev/null > gpurun_out/r2_bench_ops.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ops.json'))
for k,v in d.items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','ms_reference','speedup','ms_with_canonical','frac','ms_reference_kernels')})"
