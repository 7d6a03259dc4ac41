"""pruned FPS: 16 warps x NS slots (default) against 32 warps x NS/2; indices must agree"""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import synth
from pointrcnn_b200 import _cabi
from pointrcnn_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
for (B, N, M, prune) in [(16, 16384, 4096, 1), (16, 8192, 2048, 1), (16, 4096, 1024, 2), (2, 16384, 4096, 1)]:
    x = torch.from_numpy(synth.u_kitti(B, N, 5)[..., :3].copy()).to(dev)
    base = None
    for thr in (0, 1024):
        with _cabi.options(fps_prune=prune, fps_threads=thr):
            idx = pu.furthest_point_sample(x, M); torch.cuda.synchronize()
            if base is None: base = idx
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); pu.furthest_point_sample(x, M); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
        ms = sorted(ts)[2]
        print(json.dumps(dict(B=B, N=N, M=M, threads=thr or 512, ms=round(ms, 4), ns_per_round=round(ms * 1e6 / (M - 1)), same=bool(torch.equal(idx, base)))), flush=True)
