"""oracle/make_golden_stage_losses.py -- golden vectors for the stage losses (lib/net/train_functions.py:55-209: get_rpn_loss,
get_rcnn_loss).  The two functions are closures inside `model_joint_fn_decorator`; their SOURCE is taken from the reference file
unchanged (ast), compiled with the globals they close over (cfg, loss_utils, MEAN_SIZE on the CPU) and run on seeded inputs.
Writes tests/golden/stage_losses.npz.  TEST INFRASTRUCTURE ONLY.   python oracle/make_golden_stage_losses.py"""
import ast
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)


def rpn_inputs(seed=400, B=2, N=3000, C=76):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(B, N, 1, generator=g)
    reg = torch.randn(B, N, C, generator=g) * 0.5
    lab = (torch.rand(B, N, generator=g) < 0.05).long()
    lab[torch.rand(B, N, generator=g) < 0.02] = -1
    rl = torch.randn(B, N, 7, generator=g) * torch.tensor([1.2, 0.3, 1.2, 0.15, 0.15, 0.4, 1.5]) + torch.tensor([0, 0, 0, 1.5, 1.6, 3.9, 0.0])
    return cls, reg, lab, rl


def rcnn_inputs(seed=500, R=256, C=46):
    g = torch.Generator().manual_seed(seed)
    cls = torch.randn(R, 1, generator=g)
    reg = torch.randn(R, C, generator=g) * 0.5
    lab = torch.randint(-1, 2, (R,), generator=g)
    valid = ((lab == 1) & (torch.rand(R, generator=g) < 0.8)).long()
    roi = torch.randn(R, 7, generator=g) * torch.tensor([10, 1, 20, 0.15, 0.15, 0.4, 1.5]) + torch.tensor([0, 1.6, 30, 1.5, 1.6, 3.9, 0.0])
    gt = torch.randn(R, 7, generator=g) * torch.tensor([0.6, 0.2, 0.6, 0.15, 0.15, 0.4, 0.5]) + torch.tensor([0, 0, 0, 1.5, 1.6, 3.9, 0.0])
    return cls, reg, lab, valid, roi, gt


def main():
    from pointrcnn_b200 import dropin
    dropin._install_compat()
    torch.Tensor.get_device = lambda self: "cpu"
    torch.cuda.FloatTensor = lambda *a: torch.FloatTensor(*a)      # loss_utils allocates its one-hots with torch.cuda.FloatTensor
    from lib.config import cfg, cfg_from_file
    cfg_from_file(os.path.join(REF, "tools", "cfgs", "default.yaml"))      # the configuration the mirrors' defaults follow
    import lib.utils.loss_utils as loss_utils
    src = open(os.path.join(REF, "lib", "net", "train_functions.py")).read()
    outer = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "model_joint_fn_decorator"][0]
    fns = [n for n in outer.body if isinstance(n, ast.FunctionDef) and n.name in ("get_rpn_loss", "get_rcnn_loss")]
    ns = dict(torch=torch, nn=nn, F=F, loss_utils=loss_utils, cfg=cfg, MEAN_SIZE=torch.from_numpy(cfg.CLS_MEAN_SIZE[0]))
    exec(compile(ast.Module(body=fns, type_ignores=[]), "train_functions.py", "exec"), ns)
    out = {}
    # ---- RPN loss, the three classification variants
    cls, reg, lab, rl = rpn_inputs()
    for name, func in (("SigmoidFocalLoss", loss_utils.SigmoidFocalClassificationLoss(alpha=cfg.RPN.FOCAL_ALPHA[0], gamma=cfg.RPN.FOCAL_GAMMA)),
                       ("DiceLoss", loss_utils.DiceLoss(ignore_target=-1)), ("BinaryCrossEntropy", F.binary_cross_entropy)):
        cfg.RPN.LOSS_CLS = name
        model = types.SimpleNamespace(rpn=types.SimpleNamespace(rpn_cls_loss_func=func))
        tb = {}
        loss = ns["get_rpn_loss"](model, cls.clone(), reg.clone(), lab.clone(), rl.clone(), tb)
        out["rpn_%s" % name] = np.array([float(loss), tb["rpn_loss_cls"], tb["rpn_loss_reg"], tb["rpn_loss_loc"], tb["rpn_loss_angle"], tb["rpn_loss_size"]])
    # ---- RCNN loss
    rc, rr, rlab, valid, roi, gt = rcnn_inputs()
    ret = dict(rcnn_cls=rc, rcnn_reg=rr, cls_label=rlab, reg_valid_mask=valid, roi_boxes3d=roi, gt_of_rois=gt, pts_input=torch.zeros(rc.shape[0], 4, 3))
    for name, func in (("BinaryCrossEntropy", F.binary_cross_entropy),
                       ("SigmoidFocalLoss", loss_utils.SigmoidFocalClassificationLoss(alpha=cfg.RCNN.FOCAL_ALPHA[0], gamma=cfg.RCNN.FOCAL_GAMMA))):
        for on_roi in (False, True):
            cfg.RCNN.LOSS_CLS, cfg.RCNN.SIZE_RES_ON_ROI = name, on_roi
            model = types.SimpleNamespace(rcnn_net=types.SimpleNamespace(cls_loss_func=func))
            tb = {}
            inp = {k: v.clone() for k, v in ret.items()}
            if name == "BinaryCrossEntropy":      # torch >= 1.x rejects the reference's -1 ("ignore") targets in F.binary_cross_entropy:
                inp["cls_label"] = inp["cls_label"].clamp(min=0)      # this variant is pinned on {0,1} labels only
            loss = ns["get_rcnn_loss"](model, inp, tb)
            out["rcnn_%s_%d" % (name, int(on_roi))] = np.array([float(loss), tb["rcnn_loss_cls"], tb["rcnn_loss_reg"], tb["rcnn_reg_fg"]])
    path = os.path.join(ROOT, "tests", "golden", "stage_losses.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.tolist() for k, v in out.items()})


if __name__ == "__main__":
    main()
