#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests10.log
tail -8 gpurun_out/r2_tests10.log | cut -c1-200
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["launch"][:44].ljust(44), d["ms"], "IA", d["issuerA"], "IB", d["issuerB"], "G", d["gather0"], "E", d["epi0"])
'
short='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["launch"][:44].ljust(44), d["ms"], "G", d["gather0"], "E", d["epi0"])
'
echo "--- default"
timeout 200 python scripts/pipe_trace.py 2>&1 | tail -30 | grep "^{\|rror" | python -c "$show" | cut -c1-420
echo "--- occ=1 ne=2 ngw=2"
timeout 200 python scripts/pipe_trace.py mlp_occ=1 mlp_ne=2 mlp_ngw=2 2>&1 | tail -30 | grep "^{\|rror" | python -c "$short" | cut -c1-300
echo "--- occ=1 ne=2 ngw=3"
timeout 200 python scripts/pipe_trace.py mlp_occ=1 mlp_ne=2 mlp_ngw=3 2>&1 | tail -30 | grep "^{\|rror" | python -c "$short" | cut -c1-300
PRB_PROF_DETAIL=1 timeout 600 python bench.py --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --profile-out gpurun_out/r2_bench_b10.json > gpurun_out/r2_bench_b10.log 2>&1 || tail -5 gpurun_out/r2_bench_b10.log
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_b10.json"))
    print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"])
    for k in d["kernels"]:
        print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac") or 0, 3))
    for c in d["chain_plans"]: print("   plan", c)
except Exception as e:
    print("bench failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mlp_pipe_kernel -o gpurun_out/r2_pipe_ncu3 -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_pipe3.log 2>&1
tail -2 gpurun_out/r2_ncu_pipe3.log
ncu -i gpurun_out/r2_pipe_ncu3.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu3_raw.csv 2>/dev/null
sz=$(stat -c %s gpurun_out/r2_pipe_ncu3.ncu-rep)
if [ "$sz" -gt 50000000 ]; then rm gpurun_out/r2_pipe_ncu3.ncu-rep; echo "report dropped ($sz bytes)"; fi
du -sh gpurun_out
