"""`iou3d_cuda` -- same exports as lib/utils/iou3d/src/iou3d.cpp:174-179."""
import ctypes

import torch

from .. import _cabi as C


def _check_boxes(t, name):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float32):
        raise RuntimeError("%s must be a contiguous CUDA float tensor" % name)  # CHECK_INPUT, iou3d.cpp:8-10


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _check_boxes(boxes_a, "boxes_a"); _check_boxes(boxes_b, "boxes_b"); _check_boxes(ans_overlap, "ans_overlap")
    with torch.cuda.device(boxes_a.device):
        C.check(C.lib().prb_boxes_overlap_bev(int(boxes_a.size(0)), C.ptr(boxes_a), int(boxes_b.size(0)), C.ptr(boxes_b),
                                              C.ptr(ans_overlap), C.stream()), "boxes_overlap_bev")
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _check_boxes(boxes_a, "boxes_a"); _check_boxes(boxes_b, "boxes_b"); _check_boxes(ans_iou, "ans_iou")
    with torch.cuda.device(boxes_a.device):
        C.check(C.lib().prb_boxes_iou_bev(int(boxes_a.size(0)), C.ptr(boxes_a), int(boxes_b.size(0)), C.ptr(boxes_b),
                                          C.ptr(ans_iou), C.stream()), "boxes_iou_bev")
    return 1


def _nms(boxes, keep, thresh, normal):
    _check_boxes(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64 or not keep.is_contiguous():
        raise RuntimeError("keep must be a contiguous CPU int64 tensor")
    n = int(boxes.size(0))
    if n == 0:
        return 0
    if boxes.dim() != 2 or boxes.size(1) != 5:
        raise RuntimeError("boxes must be (N, 5)")
    if keep.numel() < n:
        raise RuntimeError("keep must hold at least N = %d entries" % n)   # the reference would overrun it silently
    with torch.cuda.device(boxes.device):
        ws = torch.empty(C.lib().prb_nms_workspace_bytes(n), dtype=torch.uint8, device=boxes.device)
        num = ctypes.c_int(0)
        C.check(C.lib().prb_nms_host(C.ptr(boxes), n, C.c_float(thresh), int(normal), C.ptr(keep), ctypes.byref(num),
                                     C.ptr(ws), C.stream()), "nms")
    return int(num.value)


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 0)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, 1)


def nms_device(boxes, thresh, normal):
    """extension: device-resident NMS -> (keep_dev int64 (n), num_dev int32 (1)); no host sync"""
    _check_boxes(boxes, "boxes")
    n = int(boxes.size(0))
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=boxes.device)
    num = torch.zeros(1, dtype=torch.int32, device=boxes.device)
    if n:
        with torch.cuda.device(boxes.device):
            ws = torch.empty(C.lib().prb_nms_workspace_bytes(n), dtype=torch.uint8, device=boxes.device)
            C.check(C.lib().prb_nms_device(C.ptr(boxes), n, C.c_float(thresh), int(normal), C.ptr(keep), C.ptr(num),
                                           C.ptr(ws), C.stream()), "nms_device")
    return keep, num


def boxes_iou3d(boxes_a, boxes_b):
    """extension: fused boxes_iou3d_gpu.  boxes_a (M,7) x boxes_b (N,7) -> (M,N); or batched (B,M,7) x (B,N,7) -> (B,M,N)"""
    _check_boxes(boxes_a, "boxes_a"); _check_boxes(boxes_b, "boxes_b")
    batched = boxes_a.dim() == 3
    a3 = boxes_a if batched else boxes_a.unsqueeze(0)
    b3 = boxes_b if batched else boxes_b.unsqueeze(0)
    if a3.size(-1) != 7 or b3.size(-1) != 7 or a3.size(0) != b3.size(0):
        raise RuntimeError("boxes_iou3d: expected (.., 7) boxes with equal batch sizes")
    out = torch.empty((a3.size(0), a3.size(1), b3.size(1)), dtype=torch.float32, device=boxes_a.device)
    with torch.cuda.device(boxes_a.device):
        C.check(C.lib().prb_boxes_iou3d(int(a3.size(0)), int(a3.size(1)), C.ptr(a3), int(b3.size(1)), C.ptr(b3), C.ptr(out), C.stream()),
                "boxes_iou3d")
    return out if batched else out[0]


def boxes_iou3d_aligned(boxes_a, boxes_b):
    """extension: out[k] = IoU3D(boxes_a[k], boxes_b[k]) for (K,7) x (K,7)"""
    _check_boxes(boxes_a, "boxes_a"); _check_boxes(boxes_b, "boxes_b")
    if boxes_a.shape != boxes_b.shape or boxes_a.dim() != 2 or boxes_a.size(1) != 7:
        raise RuntimeError("boxes_iou3d_aligned: expected two (K, 7) tensors")
    out = torch.empty(boxes_a.size(0), dtype=torch.float32, device=boxes_a.device)
    with torch.cuda.device(boxes_a.device):
        C.check(C.lib().prb_boxes_iou3d_aligned(int(boxes_a.size(0)), C.ptr(boxes_a), C.ptr(boxes_b), C.ptr(out), C.stream()),
                "boxes_iou3d_aligned")
    return out
