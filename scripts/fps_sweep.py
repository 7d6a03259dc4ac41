"""Time the FPS kernel variants (cluster size x exchange mechanism) on the RPN level shapes; verify they agree."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import synth
from pointrcnn_b200.pointnet2 import pointnet2_utils as pu

dev = torch.device("cuda:0")
res = []
for (B, N, M) in [(16, 16384, 4096), (32, 16384, 4096), (16, 4096, 1024), (2, 16384, 4096)]:
    x = torch.from_numpy(synth.u_kitti(B, N, 5)[..., :3].copy()).to(dev)
    base = None
    for cs in (1, 2, 4, 8):
        for xchg in (0,):
            for thr in ((0, 1024) if cs == 1 else (0,)):
                os.environ["PRB_FPS_CS"] = str(cs); os.environ["PRB_FPS_XCHG"] = str(xchg); os.environ["PRB_FPS_THREADS"] = str(thr)
                try:
                    idx = pu.furthest_point_sample(x, M)
                    torch.cuda.synchronize()
                except RuntimeError as e:
                    res.append(dict(B=B, N=N, M=M, cs=cs, xchg=xchg, thr=thr, err=str(e)[:80])); continue
                if base is None: base = idx
                ok = bool(torch.equal(idx, base))
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(3): pu.furthest_point_sample(x, M)
                b.record(); torch.cuda.synchronize()
                ms = a.elapsed_time(b) / 3
                res.append(dict(B=B, N=N, M=M, cs=cs, xchg=xchg, thr=thr, ms=round(ms, 3), ns_per_round=round(ms * 1e6 / (M - 1)), same=ok))
                print(res[-1], flush=True)
json.dump(res, open("gpurun_out/fps_sweep.json", "w"), indent=1)
