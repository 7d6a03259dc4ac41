#!/bin/bash
# ncu --set full of the kernels changed after the first ops capture (NMS mask / scan, input pipeline, proofs): scripts/ops_once.py again
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:"nms_|kitti_|rpn_labels|fps_prefix|roipool3d" -o gpurun_out/r2_ops2_ncu -f python scripts/ops_once.py > gpurun_out/r2_ncu_ops2.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/r2_ncu_ops2.log
ncu -i gpurun_out/r2_ops2_ncu.ncu-rep --page raw --csv > gpurun_out/r2_ops2_ncu_raw.csv 2>/dev/null
python scripts/ncu_summary.py gpurun_out/r2_ops2_ncu_raw.csv gpurun_out/r2_ncu_ops2_summary.csv
ls -la gpurun_out/r2_ops2_ncu.ncu-rep
find gpurun_out -name r2_ops2_ncu.ncu-rep -size +30M -delete
