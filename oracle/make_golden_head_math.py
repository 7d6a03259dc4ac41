"""oracle/make_golden_head_math.py -- golden vectors for the host-side torch mirrors next to the path:
decode_bbox_target (lib/utils/bbox_transform.py:24-121) and the losses (lib/utils/loss_utils.py:7-233), produced by the
reference's OWN Python imported from /root/reference, on CPU.  Writes tests/golden/head_math.npz.
TEST INFRASTRUCTURE ONLY.   python oracle/make_golden_head_math.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
ANCHOR = [1.52563191462, 1.62856739989, 3.88311640418]
DECODE_CASES = [(1.5, 0.5, 9, True, False, 7), (3.0, 0.5, 12, False, False, 3), (1.5, 0.5, 9, True, True, 7)]   # scope, bin, heads, fine_ry, y_by_bin, roi cols
LOSS_CASES = [(3.0, 12, False, False), (1.5, 9, True, False), (1.5, 9, True, True)]                                 # scope, heads, fine_ry, y_by_bin


def reg_channels(scope, bin_size, nh, ybin):
    nloc = int(scope / bin_size) * 2
    return nloc * 4 + nh * 2 + 3 + (1 if not ybin else 2 * int(0.5 / 0.25) * 2)


def decode_inputs(i, case):
    scope, bs, nh, fine, ybin, cols = case
    g = torch.Generator().manual_seed(100 + i)
    return torch.randn(400, cols, generator=g) * 3, torch.randn(400, reg_channels(scope, bs, nh, ybin), generator=g)


def loss_inputs(i, case):
    scope, nh, fine, ybin = case
    g = torch.Generator().manual_seed(200 + i)
    pred = torch.randn(300, reg_channels(scope, 0.5, nh, ybin), generator=g)
    lab = torch.randn(300, 7, generator=g) * torch.tensor([1.5, 0.3, 1.5, 0.2, 0.2, 0.4, 2.0]) + torch.tensor([0, 0, 0, 1.5, 1.6, 3.9, 0.0])
    return pred, lab


def cls_inputs():
    g = torch.Generator().manual_seed(300)
    return torch.randn(2000, generator=g), (torch.rand(2000, generator=g) > 0.8).float(), torch.rand(2000, generator=g)


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    torch.Tensor.get_device = lambda self: "cpu"
    torch.cuda.FloatTensor = lambda *a: torch.FloatTensor(*a)      # the reference allocates its one-hots with torch.cuda.FloatTensor
    bt = _load(os.path.join(REF, "lib/utils/bbox_transform.py"), "ref_bbox_transform")
    lu = _load(os.path.join(REF, "lib/utils/loss_utils.py"), "ref_loss_utils")
    anchor = torch.tensor(ANCHOR)
    out = {}
    for i, case in enumerate(DECODE_CASES):
        scope, bs, nh, fine, ybin, cols = case
        roi, reg = decode_inputs(i, case)
        out["decode_%d" % i] = bt.decode_bbox_target(roi.clone(), reg.clone(), scope, bs, nh, anchor, True, ybin, 0.5, 0.25, fine).numpy()
    for i, case in enumerate(LOSS_CASES):
        scope, nh, fine, ybin = case
        pred, lab = loss_inputs(i, case)
        loc, ang, size, _ = lu.get_reg_loss(pred, lab.clone(), scope, 0.5, nh, anchor, True, ybin, 0.5, 0.25, fine)
        out["reg_loss_%d" % i] = np.array([float(loc), float(ang), float(size)], dtype=np.float32)
    logits, tgt, w = cls_inputs()
    out["focal"] = lu.SigmoidFocalClassificationLoss(2.0, 0.25)(logits, tgt, w).numpy()
    out["dice"] = np.array([float(lu.DiceLoss()(logits, tgt))], dtype=np.float32)
    path = os.path.join(ROOT, "tests", "golden", "head_math.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
