"""every non-chain kernel family once, after warm-ups, inside a cudaProfilerStart/Stop range (target of
`ncu --profile-from-start off --set full`): the whole RPN stage of one batch (sampling incl. the proven nested levels, grid
build, ball queries, 3-NN, chain kernels, heads, proposal layer), roipool3d at the configs[3] shape, rotated / normal NMS,
the fused IoU matrix, the input pipeline and the KITTI image boxes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import numpy as np
import torch
import bench, synth
import bench_eval_e2e as E
from pointrcnn_b200.datasets import kitti_output
from pointrcnn_b200.datasets.kitti_rcnn_dataset import RPNInputPipeline
from pointrcnn_b200.ext import iou3d_cuda, roipool3d_cuda

dev = torch.device("cuda:0")
net = bench.build_model(dev)
stage = bench.build_rpn_stage(dev, net)
pc = torch.from_numpy(bench.make_scenes(0, bench.BATCH)).to(dev)
# roipool3d, BASELINE configs[3]
B, N, M, C, S = 4, 16384, 512, 130, 512
rng = np.random.default_rng(0)
xyz = synth.u_kitti(B, N, 40)
boxes = np.stack([synth.boxes3d(M, 41 + b)[0] for b in range(B)]).astype(np.float32)
for b in range(B):
    pick = rng.integers(0, N, M)
    boxes[b, :, 0], boxes[b, :, 2], boxes[b, :, 1] = xyz[b, pick, 0], xyz[b, pick, 2], xyz[b, pick, 1] + 0.8
    boxes[b, :, 3:6] += 2.0; boxes[b, :, 1] += 1.0
x, bx, f = torch.from_numpy(xyz).to(dev), torch.from_numpy(boxes).to(dev), torch.randn(B, N, C, device=dev)
pooled = torch.zeros((B, M, S, 3 + C), device=dev); empty = torch.zeros((B, M), dtype=torch.int32, device=dev)
bev_r = torch.from_numpy(synth.sorted_bev(1000, 1050)).to(dev)
bev_n = torch.from_numpy(synth.sorted_bev(6300, 6350)).to(dev)
b3 = torch.from_numpy(np.stack([synth.boxes3d(512, 60 + i)[0] for i in range(4)])).to(dev)
g3 = torch.from_numpy(np.stack([synth.boxes3d(16, 70 + i)[0] for i in range(4)])).to(dev)
scans = [dict(lidar=torch.from_numpy(E.synth_raw_scan(9000 + i, 120000)).pin_memory(), calib=E.CALIB, img_shape=E.IMG_SHAPE,
              gt_boxes3d=synth.boxes3d(8, 80 + i)[0], gt_alpha=np.zeros(8, np.float32)) for i in range(8)]
pipe = RPNInputPipeline(npoints=16384, mode="TRAIN", draw="device", device=dev)
det = torch.from_numpy(np.stack([synth.boxes3d(100, 90 + i)[0] for i in range(8)])).to(dev)


def everything():
    with torch.no_grad():
        stage(pc)
    roipool3d_cuda.forward(x, bx, f, pooled, empty, bx)
    iou3d_cuda.nms_device(bev_r, 0.3, 0)
    iou3d_cuda.nms_device(bev_n, 0.8, 1)
    iou3d_cuda.boxes_iou3d(b3, g3)
    pipe.prepare_batch(scans, seed=1)
    kitti_output.write_kitti_batch(range(8), [E.CALIB] * 8, [E.IMG_SHAPE] * 8, det, torch.rand(8, 100, device=dev), torch.ones(8, 100, device=dev))


for _ in range(3):
    everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
everything()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
