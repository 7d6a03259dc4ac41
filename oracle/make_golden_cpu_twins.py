"""oracle/make_golden_cpu_twins.py -- golden vectors for the CPU twins.

Runs the reference's OWN code in this container: lib/utils/roipool3d/roipool3d_utils.py (imported unchanged from
/root/reference) on top of the reference's own extension module -- roipool3d.cpp compiled unmodified by
`make -C oracle ref_twins` (oracle/_ref/ref_roipool3d_cuda.so) -- for pts_in_boxes3d_cpu, roipool_pc_cpu and
roipool3d_cpu (roipool3d.cpp:82-195, roipool3d_utils.py:31-108), and writes tests/golden/roipool3d_cpu_twins.npz.
TEST INFRASTRUCTURE ONLY.  Usage (container, no GPU needed):  python oracle/make_golden_cpu_twins.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import synth  # noqa: E402

REF = "/root/reference"
CASES = [("a", 4096, 24, 16, 64, 1.0, 31), ("b", 2000, 9, 3, 512, 0.2, 32), ("c", 300, 5, 8, 16, 0.0, 33)]   # name, N, M, C, S, extra, seed


def scene(N, M, C, seed):
    """points around the boxes (so that boxes are neither all empty nor all full) + far boxes that stay empty"""
    rng = np.random.default_rng(seed)
    boxes, _ = synth.boxes3d(M, seed, centres=max(2, M // 3))
    boxes[-1, 0] += 500.0                                   # an empty box
    which = rng.integers(0, M - 1, N)
    pts = boxes[which, :3] + rng.normal(0, 1.2, (N, 3))
    pts[:, 1] -= boxes[which, 3] / 2                        # box y is the bottom centre
    feat = rng.standard_normal((N, C))
    return pts.astype(np.float32), boxes.astype(np.float32), feat.astype(np.float32)


def load_reference_utils():
    so = os.path.join(ROOT, "oracle", "_ref", "ref_roipool3d_cuda.so")
    spec = importlib.util.spec_from_file_location("ref_roipool3d_cuda", so)
    ext = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ext)
    sys.modules["roipool3d_cuda"] = ext                      # the name roipool3d_utils.py imports
    sys.path.insert(0, REF)
    for name in ("lib", "lib.utils"):                        # plain namespace packages: avoid lib/__init__ side effects
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REF, *name.split("."))]
            sys.modules[name] = m
    import lib.utils.roipool3d.roipool3d_utils as ru          # the reference's own Python, unchanged
    return ru


def main():
    ru = load_reference_utils()
    out = {}
    for name, N, M, C, S, extra, seed in CASES:
        pts, boxes, feat = scene(N, M, C, seed)
        masks = ru.pts_in_boxes3d_cpu(torch.from_numpy(pts), torch.from_numpy(boxes))
        out[name + "_mask"] = np.stack([m.numpy() for m in masks])
        pp, pf, pe = ru.roipool_pc_cpu(torch.from_numpy(pts), torch.from_numpy(feat), torch.from_numpy(boxes), S)
        out[name + "_pc_pts"], out[name + "_pc_feat"], out[name + "_pc_empty"] = pp.numpy(), pf.numpy(), pe.numpy()
        ex = feat[:, :2].copy()
        a, b, e = ru.roipool3d_cpu(boxes, pts, feat, ex, extra, sampled_pt_num=S, canonical_transform=False)
        out[name + "_rp_input"], out[name + "_rp_feat"], out[name + "_rp_empty"] = a, b, e
        keep = np.nonzero(e == 0)[0]                         # the canonical branch is only defined for non-empty boxes
        a2, b2 = ru.roipool3d_cpu(boxes[keep], pts, feat, ex, extra, sampled_pt_num=S, canonical_transform=True)
        out[name + "_rpc_input"], out[name + "_rpc_feat"], out[name + "_rpc_keep"] = a2, b2, keep
        print(name, "inside counts", out[name + "_mask"].sum(1)[:8], "empty", int(pe.sum()))
    path = os.path.join(ROOT, "tests", "golden", "roipool3d_cpu_twins.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
