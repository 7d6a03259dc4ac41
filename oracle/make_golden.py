"""oracle/make_golden.py -- run the reference's OWN CUDA kernels (oracle/_ref) on seeded synthetic inputs and
write their outputs as golden vectors.  Must run on a GPU box (gpurun); inputs come from tests/synth.py seeds,
so only outputs are stored.  Usage:  python oracle/make_golden.py gpurun_out/golden
The resulting .npz files are committed under tests/golden/ and pin the CPU oracle (tests/test_cpu_golden.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402
from oracle import refgpu as R  # noqa: E402

GOLDEN_CASES = {
    "fps": [(2, 4096, 512, "kitti", 1), (3, 512, 128, "dup", 2), (2, 300, 50, "kitti", 3), (1, 16384, 1024, "dup", 4),
            (4, 128, 32, "dup", 5)],
    "ball_query": [(2, 4096, 256, "cube", 0.1, 32, 6), (2, 2048, 128, "kitti", 2.0, 16, 7), (2, 512, 128, "dup", 0.2, 64, 8)],
    "three_nn": [(2, 1024, 256, "kitti", 9), (2, 256, 64, "dup", 10), (1, 50, 2, "cube", 11)],
    "nms": [(100, 0.1, 0, 12), (1000, 0.3, 0, 13), (2700, 0.8, 1, 14), (6300, 0.85, 1, 15), (65, 0.5, 0, 16)],
}


def cloud(kind, B, N, seed):
    return {"kitti": synth.u_kitti, "cube": synth.u_cube, "dup": synth.dup_cloud}[kind](B, N, seed)


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = {}
    for i, (B, N, M, kind, seed) in enumerate(GOLDEN_CASES["fps"]):
        idx, temp = R.fps(T(cloud(kind, B, N, seed)), M, return_temp=True)
        out["fps_idx_%d" % i] = idx.cpu().numpy()
        out["fps_temp_%d" % i] = temp.cpu().numpy()
    for i, (B, N, M, kind, r, ns, seed) in enumerate(GOLDEN_CASES["ball_query"]):
        xyz = cloud(kind, B, N, seed)
        x = T(xyz)
        fidx = R.fps(x, M).cpu().numpy()
        new_xyz = np.stack([xyz[b][fidx[b]] for b in range(B)])
        out["bq_idx_%d" % i] = R.ball_query(r, ns, x, T(new_xyz)).cpu().numpy()
    for i, (B, n, m, kind, seed) in enumerate(GOLDEN_CASES["three_nn"]):
        unknown = cloud(kind, B, n, seed)
        known = np.ascontiguousarray(unknown[:, ::max(1, n // m)][:, :m])
        d2, idx = R.three_nn(T(unknown), T(known))
        out["nn_d2_%d" % i] = d2.cpu().numpy()
        out["nn_idx_%d" % i] = idx.cpu().numpy()
        rng = np.random.default_rng(seed)
        feats = rng.standard_normal((B, 7, m)).astype(np.float32)
        w = rng.random((B, n, 3)).astype(np.float32)
        out["interp_%d" % i] = R.three_interpolate(T(feats), idx, T(w)).cpu().numpy()
    for i, (n, thresh, normal, seed) in enumerate(GOLDEN_CASES["nms"]):
        boxes = T(synth.sorted_bev(n, seed))
        out["nms_keep_%d" % i] = R.nms(boxes, thresh, bool(normal)).numpy()
        mask = R.nms_mask(boxes, thresh, bool(normal)).cpu().numpy().view(np.uint64)
        out["nms_maskrowxor_%d" % i] = np.bitwise_xor.reduce(mask, axis=1)
    a, b = synth.sorted_bev(120, 17), synth.sorted_bev(90, 18)
    b[:40] = a[10:50] + np.float32(0.02)
    out["overlap"] = R.boxes_overlap_bev(T(a), T(b)).cpu().numpy()
    out["iou_bev"] = R.boxes_iou_bev(T(a), T(b)).cpu().numpy()
    # roipool3d: flags, selected-row checksum
    xyz = synth.u_kitti(2, 4096, 19)
    boxes = np.stack([synth.boxes3d(24, 20 + bb)[0] for bb in range(2)])
    rng = np.random.default_rng(21)
    for bb in range(2):
        pick = rng.integers(0, 4096, 12)
        boxes[bb, :12, 0], boxes[bb, :12, 2], boxes[bb, :12, 1] = xyz[bb, pick, 0], xyz[bb, pick, 2], xyz[bb, pick, 1] + 0.8
        boxes[bb, :4, 3:6] *= 6.0
    feat = rng.standard_normal((2, 4096, 5)).astype(np.float32)
    pooled, empty = R.roipool3d(T(xyz), T(feat), T(boxes.astype(np.float32)), 64)
    out["roi_pooled"] = pooled.cpu().numpy()
    out["roi_empty"] = empty.cpu().numpy()
    np.savez_compressed(os.path.join(outdir, "reference_kernels.npz"), **out)
    print("wrote", len(out), "arrays to", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden")
