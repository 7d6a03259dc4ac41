#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_chain.py -m gpu -q -p no:cacheprovider --tb=short -k "roipool or chain" 2>&1 | grep -v "Warning\|warnings.warn" | tail -8 | cut -c1-250
for d in 0 1; do
echo "roipool_direct=$d"
PRB_ROIPOOL_DIRECT=$d timeout 300 python scripts/bench_ops.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin)
v=d['roipool3d_C4']; print({a: (round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms','ms_with_canonical','frac','ms_exhaustive')})"
PRB_ROIPOOL_DIRECT=$d timeout 300 python scripts/bench_rcnn_stage.py 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('rcnn roipool incl cat', round(d['ms_roipool3d_incl_feature_cat'],4), 'total', round(d['ms_total'],3))"
PRB_ROIPOOL_DIRECT=$d timeout 300 python scripts/roipool_sweep.py 2>/dev/null | grep '"parts": 1, "stage_kb": 48'
done
