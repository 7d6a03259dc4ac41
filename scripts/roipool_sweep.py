"""roipool3d pass-B sweep at the C4 shape: CTAs per box x staging area"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import synth
from pointrcnn_b200 import _cabi
from pointrcnn_b200.ext import roipool3d_cuda
dev = torch.device("cuda:0")
B, N, M, C, S = 4, 16384, 512, 130, 512
rng = np.random.default_rng(0)
xyz = synth.u_kitti(B, N, 3)
boxes = np.stack([synth.boxes3d(M, 10 + b)[0] for b in range(B)]).astype(np.float32)
for b in range(B):
    pick = rng.integers(0, N, M)
    boxes[b, :, 0], boxes[b, :, 2], boxes[b, :, 1] = xyz[b, pick, 0], xyz[b, pick, 2], xyz[b, pick, 1] + 0.8
x, bx = torch.from_numpy(xyz).to(dev), torch.from_numpy(boxes).to(dev)
f = torch.randn(B, N, C, device=dev)
pooled = torch.empty(B, M, S, 3 + C, device=dev); empty = torch.zeros(B, M, dtype=torch.int32, device=dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
def timeit(fn, it=20):
    for _ in range(3): fn()
    ts = []
    for _ in range(it):
        flush.fill_(0.0)
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b_.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b_))
    return sorted(ts)[len(ts) // 2]
ref = None
for parts in (1, 2, 3, 4):
    for kb in (16, 24, 32, 48, 64):
        with _cabi.options(roipool_parts=parts, roipool_stage_kb=kb):
            ms = timeit(lambda: roipool3d_cuda.forward(x, bx, f, pooled, empty, None, True))
            if ref is None: ref = pooled.clone()
            same = bool(torch.equal(ref, pooled))
        print(json.dumps({"parts": parts, "stage_kb": kb, "ms": round(ms, 4), "same": same}), flush=True)
