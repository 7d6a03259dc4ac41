#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_chain.py "tests/test_gpu_ops.py::test_roipool3d_utils_and_canonical" "tests/test_proposal.py::test_proposal_target_layer_matches_reference_python" -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "^  warnings\|Warning" | tail -60 > gpurun_out/r2_tests5a.log
tail -45 gpurun_out/r2_tests5a.log | cut -c1-220
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_train.py 2>&1 | tail -12 > gpurun_out/r2_tests5b.log
tail -6 gpurun_out/r2_tests5b.log | cut -c1-200
for pj in 1 0; do
  PRB_FP_PROJECT=$pj PRB_PROF_DETAIL=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-ref-cuda --no-cpu-baseline --no-train --no-rcnn --min-seconds 0.3 --profile-out gpurun_out/r2_bench_pj$pj.json > gpurun_out/r2_bench_pj$pj.log 2>&1 || tail -5 gpurun_out/r2_bench_pj$pj.log
done
python - <<'PY'
import json
for pl in (1, 0):
    try:
        d = json.load(open("gpurun_out/r2_bench_pj%d.json" % pl))
        print("fp_project", pl, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", round(d["single_batch"]["ms_per_step"], 3))
        for k in d["kernels"]:
            print("   %-70s %.4f" % (k["name"][:70], k["ms_per_step"]), round(k.get("frac", 0), 3))
    except Exception as e:
        print("fp_project", pl, "failed", e)
PY
