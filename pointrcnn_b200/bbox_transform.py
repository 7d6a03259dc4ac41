"""Bin-based box decoding in torch -- mirror of lib/utils/bbox_transform.py:24-121 (`decode_bbox_target`), the general
form the RCNN stage and tools/eval_rcnn.py:485-493 use (RoI-relative: y offset or y bins, fine heading bins, rotation
back by the RoI heading).  The RPN-stage special case runs on the device kernel `prb_decode_rpn_proposals`
(pointrcnn_b200/rpn/proposal_layer.py); this host-side torch version is the plain composition of torch ops the
reference uses, kept here so the end-to-end chain runs without the reference tree.

Layout of pred_reg (reference :41-103): [x bins | z bins | (x res | z res) | (y bins | y res) or y offset | ry bins | ry res | hwl res]
"""
import math

import torch


def _bin_and_residual(reg, lo, nbin, bin_size, with_res, res_lo=None):
    """argmax bin over reg[:, lo:lo+nbin] -> bin index and (optionally) the residual gathered at that bin, times bin_size"""
    b = torch.argmax(reg[:, lo:lo + nbin], dim=1)
    res = None
    if with_res:
        res = torch.gather(reg[:, res_lo:res_lo + nbin], 1, b.unsqueeze(1)).squeeze(1) * bin_size
    return b, res


def rotate_xz_(box, angle):
    """in place: box[:, [0, 2]] <- [x z] @ [[cos, -sin], [sin, cos]]^T  (kitti_utils.rotate_pc_along_y_torch, bbox_transform.py:5-21)"""
    c, s = torch.cos(angle), torch.sin(angle)
    x, z = box[:, 0].clone(), box[:, 2].clone()
    box[:, 0] = x * c - z * s
    box[:, 2] = x * s + z * c
    return box


def decode_bbox_target(roi_box3d, pred_reg, loc_scope, loc_bin_size, num_head_bin, anchor_size, get_xz_fine=True,
                       get_y_by_bin=False, loc_y_scope=0.5, loc_y_bin_size=0.25, get_ry_fine=False):
    """roi_box3d (N,3) xyz or (N,7) boxes, pred_reg (N,C) -> (N,7) [x, y, z, h, w, l, ry]"""
    anchor_size = anchor_size.to(pred_reg.device)
    nloc = int(loc_scope / loc_bin_size) * 2
    ny = int(loc_y_scope / loc_y_bin_size) * 2
    off = 2 * nloc
    xb, xr = _bin_and_residual(pred_reg, 0, nloc, loc_bin_size, get_xz_fine, 2 * nloc)
    zb, zr = _bin_and_residual(pred_reg, nloc, nloc, loc_bin_size, get_xz_fine, 3 * nloc)
    pos_x = xb.float() * loc_bin_size + loc_bin_size / 2 - loc_scope
    pos_z = zb.float() * loc_bin_size + loc_bin_size / 2 - loc_scope
    if get_xz_fine:
        pos_x = pos_x + xr
        pos_z = pos_z + zr
        off = 4 * nloc
    if get_y_by_bin:
        yb, yr = _bin_and_residual(pred_reg, off, ny, loc_y_bin_size, True, off + ny)
        pos_y = yb.float() * loc_y_bin_size + loc_y_bin_size / 2 - loc_y_scope + yr + roi_box3d[:, 1]
        off += 2 * ny
    else:
        pos_y = roi_box3d[:, 1] + pred_reg[:, off]
        off += 1
    rb = torch.argmax(pred_reg[:, off:off + num_head_bin], dim=1)
    rres = torch.gather(pred_reg[:, off + num_head_bin:off + 2 * num_head_bin], 1, rb.unsqueeze(1)).squeeze(1)
    if get_ry_fine:
        per = (math.pi / 2) / num_head_bin
        ry = (rb.float() * per + per / 2) + rres * (per / 2) - math.pi / 4
    else:
        per = (2 * math.pi) / num_head_bin
        ry = (rb.float() * per + rres * (per / 2)) % (2 * math.pi)
        ry = torch.where(ry > math.pi, ry - 2 * math.pi, ry)
    off += 2 * num_head_bin
    assert off + 3 == pred_reg.shape[1], "pred_reg has %d channels, the layout needs %d" % (pred_reg.shape[1], off + 3)
    hwl = pred_reg[:, off:off + 3] * anchor_size + anchor_size
    out = torch.cat((pos_x.view(-1, 1), pos_y.view(-1, 1), pos_z.view(-1, 1), hwl, ry.view(-1, 1)), dim=1)
    if roi_box3d.shape[1] == 7:                       # RoI-relative prediction: rotate back by the RoI heading
        roi_ry = roi_box3d[:, 6]
        rotate_xz_(out, -roi_ry)
        out[:, 6] = out[:, 6] + roi_ry
    out[:, 0] = out[:, 0] + roi_box3d[:, 0]
    out[:, 2] = out[:, 2] + roi_box3d[:, 2]
    return out
