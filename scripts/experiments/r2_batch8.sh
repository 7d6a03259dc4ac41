#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -40 > gpurun_out/r2_tests8.log
tail -8 gpurun_out/r2_tests8.log | cut -c1-200
timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage8.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
head -c 1500 gpurun_out/r2_rcnn_stage8.json; echo
show='
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d["launch"][:44].ljust(44), d["ms"], "IA", d["issuerA"], "IB", d["issuerB"], "G", d["gather0"], "E", d["epi0"])
'
echo "--- occ=1 ne=2 ngw=2"
timeout 200 python scripts/pipe_trace.py mlp_occ=1 mlp_ne=2 mlp_ngw=2 2>&1 | tail -30 | grep "^{\|rror" | python -c "$show" | cut -c1-420
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -c 60 -o gpurun_out/r2_pipe_ncu2 -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_pipe2.log 2>&1
tail -2 gpurun_out/r2_ncu_pipe2.log
ncu -i gpurun_out/r2_pipe_ncu2.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu2_raw.csv 2>/dev/null
ls -la gpurun_out/r2_pipe_ncu2.ncu-rep gpurun_out/r2_pipe_ncu2_raw.csv
