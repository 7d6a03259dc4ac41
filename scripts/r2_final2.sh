#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --tb=short 2>&1 | grep -v "Warning\|warnings.warn" | tail -15 > gpurun_out/r2_tests_final.log
tail -5 gpurun_out/r2_tests_final.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 1200 python bench.py --profile-out gpurun_out/r2_bench_final.json > gpurun_out/r2_bench_final.log 2>&1 || tail -5 gpurun_out/r2_bench_final.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2_bench_final.json"))
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "single", d["single_batch"]["ms_per_step"], d["single_batch"]["ms_per_step_planned"], "launches", d["gpu_launches"])
print("vs_ref_cuda", d["vs_ref_cuda"]); print("roofline frac", d["roofline"]["frac"], "traffic", d["roofline"]["traffic"])
print("train", d["train_step"]["ms_per_step"], "rcnn", d["rcnn_stage"]["ms_total"], d["rcnn_stage"]["ms_roipool3d_incl_feature_cat"], d["rcnn_stage"]["rcnn_net"]["frac"])
for k in d["kernels"][:6]: print("  ", k["name"][:60], round(k["ms_per_step"],4), k.get("frac"))
PY
timeout 300 python scripts/bench_ops.py 2>/dev/null > gpurun_out/r2_bench_ops_final.json; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_ops_final.json'))
for k,v in d.items(): print(k, {a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','ms_reference','speedup','ms_with_canonical','frac','ms_reference_kernels','ms_exhaustive')})"
