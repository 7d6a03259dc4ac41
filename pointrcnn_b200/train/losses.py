"""Losses of the RPN / RCNN training step: mirror of lib/utils/loss_utils.py:7-233 (DiceLoss,
SigmoidFocalClassificationLoss, get_reg_loss) and of get_rpn_loss in lib/net/train_functions.py:55-120.

Same arithmetic as the reference (the same torch ops in the same order per term); what differs is bookkeeping: the
reference calls `.item()` on every partial loss (15+ host synchronisations per step, `loss_utils.py:127-128,...`,
`train_functions.py:77-78,115-118`) -- here every term stays a device tensor and the caller reads what it logs once.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class DiceLoss(nn.Module):
    def __init__(self, ignore_target=-1):
        super().__init__()
        self.ignore_target = ignore_target

    def forward(self, input, target):
        """input (N) logits, target (N) in {0,1} (ignore_target rows are masked)"""
        p = torch.sigmoid(input.view(-1))
        t = target.float().view(-1)
        mask = (t != self.ignore_target).float()
        return 1.0 - (torch.min(p, t) * mask).sum() / torch.clamp((torch.max(p, t) * mask).sum(), min=1.0)


def _sigmoid_cross_entropy_with_logits(logits, labels):
    loss = torch.clamp(logits, min=0) - logits * labels.type_as(logits)
    return loss + torch.log1p(torch.exp(-torch.abs(logits)))


class SigmoidFocalClassificationLoss(nn.Module):
    """sigmoid focal cross entropy (loss_utils.py:25-76)"""

    def __init__(self, gamma=2.0, alpha=0.25):
        super().__init__()
        self._alpha, self._gamma = alpha, gamma

    def forward(self, prediction_tensor, target_tensor, weights):
        ce = _sigmoid_cross_entropy_with_logits(labels=target_tensor, logits=prediction_tensor)
        prob = torch.sigmoid(prediction_tensor)
        p_t = (target_tensor * prob) + ((1 - target_tensor) * (1 - prob))
        mod = torch.pow(1.0 - p_t, self._gamma) if self._gamma else 1.0
        alpha_w = (target_tensor * self._alpha + (1 - target_tensor) * (1 - self._alpha)) if self._alpha is not None else 1.0
        return mod * alpha_w * ce * weights


def _bin_loss(pred, lo, nbin, shift, bin_size, with_res, res_lo):
    """cross entropy over the bins of `shift` (+ smooth-L1 on the residual inside the target bin); returns (bin_loss, res_loss, bin_label)"""
    bin_label = (shift / bin_size).floor().long()
    loss_bin = F.cross_entropy(pred[:, lo:lo + nbin], bin_label)
    loss_res = None
    if with_res:
        res_label = (shift - (bin_label.float() * bin_size + bin_size / 2)) / bin_size
        onehot = torch.zeros((bin_label.size(0), nbin), dtype=pred.dtype, device=pred.device).scatter_(1, bin_label.view(-1, 1), 1)
        loss_res = F.smooth_l1_loss((pred[:, res_lo:res_lo + nbin] * onehot).sum(dim=1), res_label)
    return loss_bin, loss_res, bin_label


def get_reg_loss(pred_reg, reg_label, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_xz_fine=True, get_y_by_bin=False,
                 loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=False):
    """bin-based box regression loss (loss_utils.py:87-233).  pred_reg (N,C), reg_label (N,7) [dx,dy,dz,h,w,l,ry] ->
    loc_loss, angle_loss, size_loss, dict of the partial terms (device tensors)"""
    nloc = int(loc_scope / loc_bin_size) * 2
    ny = int(loc_y_scope / loc_y_bin_size) * 2
    terms = {}
    x_shift = torch.clamp(reg_label[:, 0] + loc_scope, 0, loc_scope * 2 - 1e-3)
    z_shift = torch.clamp(reg_label[:, 2] + loc_scope, 0, loc_scope * 2 - 1e-3)
    lxb, lxr, _ = _bin_loss(pred_reg, 0, nloc, x_shift, loc_bin_size, get_xz_fine, 2 * nloc)
    lzb, lzr, _ = _bin_loss(pred_reg, nloc, nloc, z_shift, loc_bin_size, get_xz_fine, 3 * nloc)
    terms["loss_x_bin"], terms["loss_z_bin"] = lxb, lzb
    loc_loss = lxb + lzb
    off = 2 * nloc
    if get_xz_fine:
        terms["loss_x_res"], terms["loss_z_res"] = lxr, lzr
        loc_loss = loc_loss + lxr + lzr
        off = 4 * nloc
    y_label = reg_label[:, 1]
    if get_y_by_bin:
        y_shift = torch.clamp(y_label + loc_y_scope, 0, loc_y_scope * 2 - 1e-3)
        lyb, lyr, _ = _bin_loss(pred_reg, off, ny, y_shift, loc_y_bin_size, True, off + ny)
        terms["loss_y_bin"], terms["loss_y_res"] = lyb, lyr
        loc_loss = loc_loss + lyb + lyr
        off += 2 * ny
    else:
        ly = F.smooth_l1_loss(pred_reg[:, off:off + 1].sum(dim=1), y_label)
        terms["loss_y_offset"] = ly
        loc_loss = loc_loss + ly
        off += 1
    ry_label = reg_label[:, 6]
    if get_ry_fine:
        per = (np.pi / 2) / num_head_bin
        ry = ry_label % (2 * np.pi)
        opposite = (ry > np.pi * 0.5) & (ry < np.pi * 1.5)
        ry = torch.where(opposite, (ry + np.pi) % (2 * np.pi), ry)
        shift_angle = torch.clamp((ry + np.pi * 0.5) % (2 * np.pi) - np.pi * 0.25, min=1e-3, max=np.pi * 0.5 - 1e-3)
    else:
        per = (2 * np.pi) / num_head_bin
        shift_angle = ((ry_label % (2 * np.pi)) + per / 2) % (2 * np.pi)
    ry_bin = (shift_angle / per).floor().long()
    ry_res = (shift_angle - (ry_bin.float() * per + per / 2)) / (per / 2)
    onehot = torch.zeros((ry_bin.size(0), num_head_bin), dtype=pred_reg.dtype, device=pred_reg.device).scatter_(1, ry_bin.view(-1, 1), 1)
    l_rb = F.cross_entropy(pred_reg[:, off:off + num_head_bin], ry_bin)
    l_rr = F.smooth_l1_loss((pred_reg[:, off + num_head_bin:off + 2 * num_head_bin] * onehot).sum(dim=1), ry_res)
    terms["loss_ry_bin"], terms["loss_ry_res"] = l_rb, l_rr
    angle_loss = l_rb + l_rr
    off += 2 * num_head_bin
    assert pred_reg.shape[1] == off + 3, "%d vs %d" % (pred_reg.shape[1], off + 3)
    size_loss = F.smooth_l1_loss(pred_reg[:, off:off + 3], (reg_label[:, 3:6] - anchor_size) / anchor_size)
    terms.update(loss_loc=loc_loss, loss_angle=angle_loss, loss_size=size_loss)
    return loc_loss, angle_loss, size_loss, terms


def rpn_loss(rpn_cls, rpn_reg, rpn_cls_label, rpn_reg_label, mean_size, loss_cls="SigmoidFocalLoss", focal_alpha=0.25, focal_gamma=2.0,
             fg_weight=15.0, loc_scope=3.0, loc_bin_size=0.5, num_head_bin=12, loc_xz_fine=True, loss_weight=(1.0, 1.0)):
    """get_rpn_loss (train_functions.py:55-120).  rpn_cls (B,N,1), rpn_reg (B,N,C), rpn_cls_label (B,N) in {-1,0,1},
    rpn_reg_label (B,N,7) -> total loss, dict of device tensors.  The foreground rows are selected with a boolean mask like
    the reference (one data-dependent shape, no .item())."""
    cls_label = rpn_cls_label.view(-1)
    cls_flat = rpn_cls.view(-1)
    fg_mask = cls_label > 0
    if loss_cls == "DiceLoss":
        loss_c = DiceLoss()(rpn_cls, cls_label)
    elif loss_cls == "SigmoidFocalLoss":
        target = fg_mask.float()
        pos, neg = target, (cls_label == 0).float()
        w = (pos + neg) / torch.clamp(pos.sum(), min=1.0)
        loss_c = SigmoidFocalClassificationLoss(alpha=focal_alpha, gamma=focal_gamma)(cls_flat, target, w).sum()
    elif loss_cls == "BinaryCrossEntropy":
        weight = torch.where(fg_mask, torch.full_like(cls_flat, fg_weight), torch.ones_like(cls_flat))
        bl = F.binary_cross_entropy(torch.sigmoid(cls_flat), fg_mask.float(), weight=weight, reduction="none")
        valid = (cls_label >= 0).float()
        loss_c = (bl * valid).sum() / torch.clamp(valid.sum(), min=1.0)
    else:
        raise NotImplementedError(loss_cls)
    npts = rpn_reg.size(0) * rpn_reg.size(1)
    reg_fg = rpn_reg.view(npts, -1)[fg_mask]
    terms = {}
    if reg_fg.shape[0] != 0:
        loc, ang, size, terms = get_reg_loss(reg_fg, rpn_reg_label.view(npts, 7)[fg_mask], loc_scope=loc_scope, loc_bin_size=loc_bin_size,
                                             num_head_bin=num_head_bin, anchor_size=mean_size, get_xz_fine=loc_xz_fine,
                                             get_y_by_bin=False, get_ry_fine=False)
        loss_r = loc + ang + 3 * size             # "consistent with old codes" (train_functions.py:107)
    else:
        loss_r = loss_c * 0
    total = loss_c * loss_weight[0] + loss_r * loss_weight[1]
    terms.update(rpn_loss_cls=loss_c, rpn_loss_reg=loss_r, rpn_loss=total)
    return total, terms


def rcnn_loss(rcnn_cls, rcnn_reg, cls_label, reg_valid_mask, roi_boxes3d, gt_of_rois, mean_size, loss_cls="BinaryCrossEntropy",
              focal_alpha=0.25, focal_gamma=2.0, size_res_on_roi=False, loc_scope=1.5, loc_bin_size=0.5, num_head_bin=9,
              loc_y_by_bin=False, loc_y_scope=0.5, loc_y_bin_size=0.25, ce_weight=None):
    """get_rcnn_loss (train_functions.py:121-209).  rcnn_cls (R,1|K), rcnn_reg (R,C), cls_label (R) in {-1,0,1},
    reg_valid_mask (R), roi_boxes3d (R,7), gt_of_rois (R,7) canonical -> total loss, dict of device tensors (no .item())."""
    R = rcnn_reg.shape[0]
    label = cls_label.float().view(-1)
    terms = {}
    if loss_cls == "SigmoidFocalLoss":
        flat = rcnn_cls.view(-1)
        pos, neg = (label > 0).float(), (label == 0).float()
        w = (pos + neg) / torch.clamp(pos.sum(), min=1.0)
        per = SigmoidFocalClassificationLoss(alpha=focal_alpha, gamma=focal_gamma)(flat, pos, w)
        terms["rpn_loss_cls_pos"], terms["rpn_loss_cls_neg"] = (per * pos).sum(), (per * neg).sum()     # the reference's key names
        loss_c = per.sum()
    elif loss_cls == "BinaryCrossEntropy":
        flat = rcnn_cls.view(-1)
        # the reference passes the -1 ("ignore") labels straight into F.binary_cross_entropy and masks the rows afterwards; current
        # torch rejects targets outside [0,1], so the ignored rows get a dummy target here (they are masked out all the same)
        bl = F.binary_cross_entropy(torch.sigmoid(flat), label.clamp(min=0.0), reduction="none")
        valid = (label >= 0).float()
        loss_c = (bl * valid).sum() / torch.clamp(valid.sum(), min=1.0)
    elif loss_cls == "CrossEntropy":
        logits = rcnn_cls.view(R, -1)
        valid = (label >= 0).float()
        bl = F.cross_entropy(logits, label.long(), weight=ce_weight, ignore_index=-1, reduction="none")
        loss_c = (bl * valid).sum() / torch.clamp(valid.sum(), min=1.0)      # (the reference's .mean(dim=1) of a 1-D loss is a no-op for K classes)
    else:
        raise NotImplementedError(loss_cls)
    fg = reg_valid_mask.view(-1) > 0
    reg_fg = rcnn_reg.view(R, -1)[fg]
    if reg_fg.shape[0] != 0:
        anchor = roi_boxes3d.view(R, 7)[:, 3:6][fg] if size_res_on_roi else mean_size
        loc, ang, size, reg_terms = get_reg_loss(reg_fg, gt_of_rois.view(R, 7)[fg], loc_scope=loc_scope, loc_bin_size=loc_bin_size,
                                                 num_head_bin=num_head_bin, anchor_size=anchor, get_xz_fine=True, get_y_by_bin=loc_y_by_bin,
                                                 loc_y_scope=loc_y_scope, loc_y_bin_size=loc_y_bin_size, get_ry_fine=True)
        terms.update(reg_terms)
        loss_r = loc + ang + 3 * size              # "consistent with old codes" (train_functions.py:193)
    else:
        loss_r = loss_c * 0
    total = loss_c + loss_r
    terms.update(rcnn_loss_cls=loss_c, rcnn_loss_reg=loss_r, rcnn_loss=total)
    return total, terms
