"""nested sampling levels of the RPN encoder: sampling kernels vs the proven-prefix shortcut, and the single-batch forward
with / without it (CUDA events, L2-resident inputs are fine here: the levels are latency chains)"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from pointrcnn_b200 import config
from pointrcnn_b200.pointnet2 import pointnet2_utils as pu

dev = torch.device("cuda:0")
out = {}


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for B in (16, 2):
    pc = torch.from_numpy(bench.make_scenes(0, B)).to(dev)
    xyz = pc[..., :3].contiguous()
    _, l1 = pu.furthest_point_sample_xyz(xyz, 4096)
    res = {}
    for name, ordered in (("sampled", False), ("proven", True)):
        def chain():
            x = l1
            for m in (1024, 256, 64):
                _, x = pu.furthest_point_sample_xyz(x, m, ordered=ordered)
        res[name + "_levels_2_4_ms"] = timed(chain)
        for n_, m_, src in ((4096, 1024, l1),):
            res[name + "_level_2_ms"] = timed(lambda: pu.furthest_point_sample_xyz(src, m_, ordered=ordered))
    res["level_1_ms"] = timed(lambda: pu.furthest_point_sample_xyz(xyz, 4096), reps=10)
    net = bench.build_model(dev)
    with torch.no_grad():
        for name, flag in (("forward_sampled_ms", False), ("forward_proven_ms", True)):
            with config.override(fps_ordered=flag):
                res[name] = timed(lambda: net(pc), reps=15, warm=4)
    out["B%d" % B] = res
    print(B, json.dumps(res))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_fps_ordered.json"), "w"), indent=1)
