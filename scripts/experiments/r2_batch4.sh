#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -30 > gpurun_out/r2_tests4.log
tail -8 gpurun_out/r2_tests4.log
for br in 64 128; do
  PRB_MLP_BROWS=$br PRB_PROF_DETAIL=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline --min-seconds 0.3 --profile-out gpurun_out/r2_bench_br$br.json > gpurun_out/r2_bench_br$br.log 2>&1 || tail -5 gpurun_out/r2_bench_br$br.log
done
python - <<'PY'
import json
for pl in (64, 128):
    try:
        d = json.load(open("gpurun_out/r2_bench_br%d.json" % pl))
        print("brows", pl, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"], 3))
        for k in d["kernels"]:
            print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
    except Exception as e:
        print("brows", pl, "failed", e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -s 32 -c 16 -o gpurun_out/r2_lean_ncu -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_lean.log 2>&1
tail -2 gpurun_out/r2_ncu_lean.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:roipool -c 12 --csv --log-file gpurun_out/r2_roipool_launches.csv python scripts/bench_ops.py > /dev/null 2>&1
grep -i "roipool" gpurun_out/r2_roipool_launches.csv | awk -F'","' '{print $5, $NF}' | head -4
timeout 300 python scripts/bench_ops.py 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.readline()); print(json.dumps(d['roipool3d_C4']))"
