#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/bench_rcnn_stage.py > gpurun_out/r2_rcnn_stage9.json 2> gpurun_out/r2_rcnn_stage.err || tail -5 gpurun_out/r2_rcnn_stage.err
head -c 1300 gpurun_out/r2_rcnn_stage9.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mlp_pipe_kernel -s 38 -c 19 -o gpurun_out/r2_pipe_ncu2 -f python scripts/one_forward.py 3 > gpurun_out/r2_ncu_pipe2.log 2>&1
tail -2 gpurun_out/r2_ncu_pipe2.log
ncu -i gpurun_out/r2_pipe_ncu2.ncu-rep --page raw --csv > gpurun_out/r2_pipe_ncu2_raw.csv 2>/dev/null
ls -la gpurun_out/
# keep the transfer under the 64 MiB limit: source pages of three representative launches, then drop the report if it is too large
for id in 1 5 18; do
  ncu -i gpurun_out/r2_pipe_ncu2.ncu-rep --page source --csv --print-source sass --launch-skip $id --launch-count 1 > gpurun_out/r2_pipe_ncu2_src_$id.csv 2>/dev/null
done
sz=$(stat -c %s gpurun_out/r2_pipe_ncu2.ncu-rep)
if [ "$sz" -gt 45000000 ]; then rm gpurun_out/r2_pipe_ncu2.ncu-rep; echo "report dropped ($sz bytes)"; fi
du -sh gpurun_out
