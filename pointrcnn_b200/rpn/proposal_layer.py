"""Device-side RPN proposal layer: mirror of lib/rpn/proposal_layer.py (same class name, constructor and forward
signature) on top of prb_decode_rpn_proposals / prb_rpn_proposals (pointrcnn_b200/csrc/proposal.cu).

The reference decodes with ~40 small torch kernels, then loops over the scenes in Python: boolean-mask compaction per
distance range, two NMS calls per scene through the C++ extension (each copies the keep list to the host), torch.cat.
Here: one decode launch, one batched torch.sort (the same call the reference makes, so the order -- including how
ties fall -- is the reference's), one launch that selects and runs the top-k greedy NMS for every (scene, range), one
pack launch.  No host synchronisation; results equal the reference's bit for bit (tests/golden/proposal_layer.npz).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi as C


def _cfg():
    from lib.config import cfg          # the reference's config module (its tree must be importable), lib/config.py
    return cfg


def decode_rpn_proposals(xyz, rpn_reg, anchor_size, loc_scope, loc_bin_size, num_head_bin, get_xz_fine=True):
    """xyz (B,N,3), rpn_reg (B,N,C) -> (B,N,7) boxes [x, y(bottom centre), z, h, w, l, ry]:
    decode_bbox_target (lib/utils/bbox_transform.py:24-121) + `proposals[:, 1] += proposals[:, 3] / 2` (proposal_layer.py:32)"""
    C.require_cuda(xyz, rpn_reg)
    xyz, rpn_reg = xyz.contiguous().float(), rpn_reg.contiguous().float()
    B, N = xyz.shape[0], xyz.shape[1]
    out = torch.empty((B, N, 7), dtype=torch.float32, device=xyz.device)
    anchor = (ctypes.c_float * 3)(*[float(v) for v in np.asarray(anchor_size, dtype=np.float32).reshape(-1)[:3]])
    with torch.cuda.device(xyz.device):
        C.check(C.lib().prb_decode_rpn_proposals(C.c_long(B * N), int(rpn_reg.shape[-1]), C.ptr(xyz), C.ptr(rpn_reg), anchor,
                                                 C.c_float(loc_scope), C.c_float(loc_bin_size), int(num_head_bin),
                                                 int(bool(get_xz_fine)), C.ptr(out), C.stream()), "decode_rpn_proposals")
    return out


def rpn_proposals(boxes, scores, pre_nms_top_n, post_nms_top_n, nms_thresh, nms_type="normal", distance_based=True):
    """boxes (B,N,7), scores (B,N) -> (ret_bbox3d (B,post,7), ret_scores (B,post)), rows beyond the survivors are zero"""
    C.require_cuda(boxes, scores)
    boxes, scores = boxes.contiguous(), scores.contiguous().float()
    B, N = scores.shape
    _, order = torch.sort(scores, dim=1, descending=True)               # proposal_layer.py:36
    out_b = torch.empty((B, post_nms_top_n, 7), dtype=torch.float32, device=boxes.device)
    out_s = torch.empty((B, post_nms_top_n), dtype=torch.float32, device=boxes.device)
    lib = C.lib()
    wsb = lib.prb_rpn_proposals_workspace_bytes(B, int(post_nms_top_n))
    ws = torch.empty(wsb, dtype=torch.uint8, device=boxes.device)
    # score based proposals always use the rotated NMS (proposal_layer.py:136)
    normal = 1 if (nms_type == "normal" and distance_based) else 0
    if nms_type not in ("normal", "rotate"):
        raise NotImplementedError(nms_type)
    with torch.cuda.device(boxes.device):
        C.check(lib.prb_rpn_proposals(B, N, C.ptr(boxes), C.ptr(scores), C.ptr(order), int(bool(distance_based)),
                                      int(pre_nms_top_n), int(post_nms_top_n), C.c_float(nms_thresh), normal, C.ptr(out_b),
                                      C.ptr(out_s), C.ptr(ws), C.c_size_t(wsb), C.stream()), "rpn_proposals")
    return out_b, out_s


class ProposalLayer(nn.Module):
    def __init__(self, mode='TRAIN', cfg=None):
        super().__init__()
        self.mode = mode
        self.cfg = cfg if cfg is not None else _cfg()
        self.MEAN_SIZE = torch.from_numpy(np.asarray(self.cfg.CLS_MEAN_SIZE[0], dtype=np.float32))

    def forward(self, rpn_scores, rpn_reg, xyz):
        """rpn_scores (B,N), rpn_reg (B,N,C), xyz (B,N,3) -> bbox3d (B,M,7), scores (B,M)   [proposal_layer.py:15-56]"""
        cfg = self.cfg
        proposals = decode_rpn_proposals(xyz, rpn_reg, self.MEAN_SIZE.numpy(), cfg.RPN.LOC_SCOPE, cfg.RPN.LOC_BIN_SIZE,
                                         cfg.RPN.NUM_HEAD_BIN, cfg.RPN.LOC_XZ_FINE)
        m = cfg[self.mode]
        return rpn_proposals(proposals, rpn_scores, m.RPN_PRE_NMS_TOP_N, m.RPN_POST_NMS_TOP_N, m.RPN_NMS_THRESH,
                             nms_type=cfg.RPN.NMS_TYPE, distance_based=cfg.TEST.RPN_DISTANCE_BASED_PROPOSE)
