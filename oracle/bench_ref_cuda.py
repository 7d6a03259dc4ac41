"""oracle/bench_ref_cuda.py -- time the reference's own CUDA-extension pipeline on the bench workload (B-ref of
BASELINE.md): the UNMODIFIED reference kernels (oracle/_ref, rebuilt for sm_100a) called in exactly the order
pointnet2_modules.py calls them (ref:pointnet2_modules.py:19-55,127-156, pointnet2_utils.py:241-264), with the
MLP through stock torch.nn Conv2d/BatchNorm2d/ReLU/max_pool2d (cuDNN/cuBLAS, torch defaults incl. TF32 convs).

TEST / BASELINE INFRASTRUCTURE ONLY.  Usage on the GPU box:  python oracle/bench_ref_cuda.py [--steps 10]
Prints one JSON line (also the per-family event timings) -- recorded under profiles/.
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402  (workload definition: model, scenes, constants)
from oracle import refgpu as R  # noqa: E402
from pointrcnn_b200 import prof  # noqa: E402  (event timer only)


def sa_forward(mod, xyz, features):
    xyz_flipped = xyz.transpose(1, 2).contiguous()
    with prof.region("fps"):
        idx = R.fps(xyz, mod.npoint)
    with prof.region("gather"):
        new_xyz = R.gather(xyz_flipped, idx).transpose(1, 2).contiguous()
    outs = []
    for g, mlp in zip(mod.groupers, mod.mlps):
        with prof.region("ball_query"):
            bi = R.ball_query(g.radius, g.nsample, xyz, new_xyz)
        with prof.region("group"):
            xyz_trans = xyz.transpose(1, 2).contiguous()
            gx = R.group(xyz_trans, bi)
            gx -= new_xyz.transpose(1, 2).unsqueeze(-1)
            nf = torch.cat([gx, R.group(features, bi)], dim=1) if features is not None else gx
        with prof.region("sa_mlp"):
            nf = mlp(nf)
            nf = F.max_pool2d(nf, kernel_size=[1, nf.size(3)]).squeeze(-1)
        outs.append(nf)
    return new_xyz, torch.cat(outs, dim=1)


def fp_forward(mod, unknown, known, uf, kf):
    with prof.region("three_nn"):
        d2, idx = R.three_nn(unknown, known)
        dist = torch.sqrt(d2)
        recip = 1.0 / (dist + 1e-8)
        weight = recip / torch.sum(recip, dim=2, keepdim=True)
    with prof.region("interpolate"):
        interp = R.three_interpolate(kf, idx, weight)
    with prof.region("fp_mlp"):
        x = torch.cat([interp, uf], dim=1) if uf is not None else interp
        return mod.mlp(x.unsqueeze(-1)).squeeze(-1)


def backbone(net, pc):
    xyz, feats = net._break_up_pc(pc)
    l_xyz, l_f = [xyz], [feats]
    for sa in net.SA_modules:
        nx, nf = sa_forward(sa, l_xyz[-1], l_f[-1])
        l_xyz.append(nx)
        l_f.append(nf)
    for i in range(-1, -(len(net.FP_modules) + 1), -1):
        l_f[i - 1] = fp_forward(net.FP_modules[i], l_xyz[i - 1], l_xyz[i], l_f[i - 1], l_f[i])
    return l_xyz[0], l_f[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    net = bench.build_model(dev)
    pc = torch.from_numpy(bench.make_scenes(0, bench.BATCH)).to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    with torch.no_grad():
        for _ in range(a.warmup):
            backbone(net, pc)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
        prof.enable()
        for s, e in ev:
            flush.fill_(1.0)
            s.record()
            _, feats = backbone(net, pc)
            e.record()
        torch.cuda.synchronize()
        prof.disable()
    ms = sum(s.elapsed_time(e) for s, e in ev) / a.steps
    fam = {k: v[0] / a.steps for k, v in prof.collect().items()}
    print(json.dumps({"impl": "reference CUDA-extension build (oracle/_ref kernels, sm_100a, + torch cuDNN MLP)",
                      "metric": "scenes/sec RPN backbone fwd (16384 pts)", "value": bench.BATCH / (ms * 1e-3), "unit": "scenes/s",
                      "ms_per_step": ms, "steps": a.steps, "batch": bench.BATCH, "family_ms_per_step": fam,
                      "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "checksum": float(feats.double().mean())}))


if __name__ == "__main__":
    main()
