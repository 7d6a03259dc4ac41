"""oracle/refgpu.py -- torch front-end of oracle/_ref/libref_pointops.so: the reference's OWN CUDA kernels
(compiled unmodified from /root/reference by oracle/Makefile) for bit-exact parity checks and as the timed
"reference CUDA-extension build" comparator (BASELINE.md B-ref).

TEST INFRASTRUCTURE ONLY: imported by tests/, oracle/make_golden.py and bench.py's comparator leg.
The .so is built in the container (where /root/reference exists) and travels to the GPU box.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref_pointops.so")
_lib = None
c_float, c_void_p = ctypes.c_float, ctypes.c_void_p


def available():
    return os.path.exists(SO)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(SO)
    return _lib


def _p(t):
    return c_void_p(t.data_ptr())


def _st():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def fps(xyz, npoint, return_temp=False):
    B, N, _ = xyz.shape
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    lib().ref_fps(B, N, int(npoint), _p(xyz), _p(temp), _p(idx), _st())
    return (idx, temp) if return_temp else idx


def fps_raw(xyz, temp, idx):
    """the reference kernel on caller-owned temp (in/out) and idx"""
    B, N, _ = xyz.shape
    lib().ref_fps(B, N, int(idx.shape[1]), _p(xyz), _p(temp), _p(idx), _st())


def gather(features, idx):
    B, C, N = features.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=features.device)
    lib().ref_gather(B, C, N, M, _p(features), _p(idx), _p(out), _st())
    return out


def gather_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    lib().ref_gather_grad(B, C, N, M, _p(grad_out), _p(idx), _p(g), _st())
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)
    lib().ref_ball_query(B, N, M, c_float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx), _st())
    return idx


def group(features, idx):
    B, C, N = features.shape
    _, M, S = idx.shape
    out = torch.empty((B, C, M, S), dtype=torch.float32, device=features.device)
    lib().ref_group(B, C, N, M, S, _p(features), _p(idx), _p(out), _st())
    return out


def group_grad(grad_out, idx, N):
    B, C, M, S = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    lib().ref_group_grad(B, C, N, M, S, _p(grad_out), _p(idx), _p(g), _st())
    return g


def three_nn(unknown, known):
    B, N, _ = unknown.shape
    M = known.shape[1]
    dist2 = torch.empty((B, N, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=unknown.device)
    lib().ref_three_nn(B, N, M, _p(unknown), _p(known), _p(dist2), _p(idx), _st())
    return dist2, idx


def three_interpolate(features, idx, weight):
    B, C, M = features.shape
    N = idx.shape[1]
    out = torch.empty((B, C, N), dtype=torch.float32, device=features.device)
    lib().ref_three_interpolate(B, C, M, N, _p(features), _p(idx), _p(weight), _p(out), _st())
    return out


def three_interpolate_grad(grad_out, idx, weight, M):
    B, C, N = grad_out.shape
    g = torch.zeros((B, C, M), dtype=torch.float32, device=grad_out.device)
    lib().ref_three_interpolate_grad(B, C, N, M, _p(grad_out), _p(idx), _p(weight), _p(g), _st())
    return g


# iou3d / roipool3d launch on the legacy default stream like the reference; callers synchronise around them
def boxes_overlap_bev(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    torch.cuda.synchronize()
    lib().ref_boxes_overlap_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    torch.cuda.synchronize()
    return out


def boxes_iou_bev(a, b):
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    torch.cuda.synchronize()
    lib().ref_boxes_iou_bev(a.shape[0], _p(a), b.shape[0], _p(b), _p(out))
    torch.cuda.synchronize()
    return out


def nms_mask(boxes, thresh, normal=False):
    n = boxes.shape[0]
    mask = torch.zeros((n, (n + 63) // 64), dtype=torch.int64, device=boxes.device)
    torch.cuda.synchronize()
    lib().ref_nms_mask(_p(boxes), n, c_float(thresh), int(bool(normal)), _p(mask))
    torch.cuda.synchronize()
    return mask


def nms(boxes, thresh, normal=False):
    """whole reference nms_gpu: cudaMalloc + kernel + blocking D2H + host scan; returns kept positions (CPU int64)"""
    n = boxes.shape[0]
    keep = torch.zeros(max(n, 1), dtype=torch.int64)
    torch.cuda.synchronize()
    num = lib().ref_nms(_p(boxes), n, c_float(thresh), int(bool(normal)), c_void_p(keep.data_ptr()))
    return keep[:num]


def roipool3d(xyz, pts_feature, boxes3d, S=512, slow=False):
    B, N, _ = xyz.shape
    M, C = boxes3d.shape[1], pts_feature.shape[2]
    pooled = torch.zeros((B, M, S, 3 + C), dtype=torch.float32, device=xyz.device)
    empty = torch.zeros((B, M), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    fn = lib().ref_roipool3d_slow if slow else lib().ref_roipool3d
    fn(B, N, M, C, int(S), _p(xyz), _p(boxes3d), _p(pts_feature), _p(pooled), _p(empty))
    torch.cuda.synchronize()
    return pooled, empty
