"""KITTI-format result files: mirror of tools/eval_rcnn.py:69-94 (save_kitti_format) -- SURVEY.md 8(f) rank 4.

The box arithmetic (8 corners, projection through P2, clipping, the 0.8-of-the-image validity rule, alpha) is one launch
over the scene's detections (`prb_kitti_image_boxes`); the text is formatted by the library's host function
(`prb_kitti_format_detections`, one snprintf per line instead of a Python `print` per box).
"""
import ctypes
import os

import numpy as np
import torch

from .. import _cabi as C
from .kitti_rcnn_dataset import Calibration


def kitti_image_boxes(bbox3d, P2, img_shape):
    """bbox3d (N,7) CUDA float32 -> (img_boxes (N,4) float32, alpha (N) float32, valid (N) int32), all CUDA"""
    C.require_cuda(bbox3d)
    b = bbox3d.contiguous().float()
    n = b.size(0)
    dev = b.device
    img_boxes = torch.empty((n, 4), dtype=torch.float32, device=dev)
    alpha = torch.empty(n, dtype=torch.float32, device=dev)
    valid = torch.empty(n, dtype=torch.int32, device=dev)
    p2 = torch.as_tensor(np.asarray(P2, dtype=np.float32).reshape(-1)).to(dev)
    with torch.cuda.device(dev):
        C.check(C.lib().prb_kitti_image_boxes(n, C.ptr(b), C.ptr(p2), ctypes.c_float(float(img_shape[0])), ctypes.c_float(float(img_shape[1])),
                                              C.ptr(img_boxes), C.ptr(alpha), C.ptr(valid), C.stream()), "kitti_image_boxes")
    return img_boxes, alpha, valid


def format_kitti_lines(bbox3d, img_boxes, alpha, scores, valid, classes="Car"):
    """host arrays -> the file's text (str)"""
    b = np.ascontiguousarray(bbox3d, dtype=np.float32)
    ib = np.ascontiguousarray(img_boxes, dtype=np.float32)
    al = np.ascontiguousarray(alpha, dtype=np.float32)
    sc = np.ascontiguousarray(scores, dtype=np.float32).reshape(-1)
    va = np.ascontiguousarray(valid, dtype=np.int32)
    n = b.shape[0]
    lib = C.lib()
    P = ctypes.c_void_p
    args = (classes.encode(), n, P(b.ctypes.data), P(ib.ctypes.data), P(al.ctypes.data), P(sc.ctypes.data), P(va.ctypes.data))
    need = lib.prb_kitti_format_detections(*args, None, ctypes.c_size_t(0))
    buf = ctypes.create_string_buffer(need + 1)
    lib.prb_kitti_format_detections(*args, buf, ctypes.c_size_t(need + 1))
    return buf.value.decode()


def detections_to_host(bbox3d, scores, P2, img_shape):
    """device work + ONE device-to-host copy for a scene's detections: -> (bbox3d, img_boxes, alpha, scores, valid) numpy"""
    n = bbox3d.size(0)
    img_boxes, alpha, valid = kitti_image_boxes(bbox3d, P2, img_shape)
    packed = torch.cat((bbox3d.float(), img_boxes, alpha.unsqueeze(1), scores.float().view(n, 1), valid.float().unsqueeze(1)), dim=1)
    h = packed.cpu().numpy()
    return h[:, 0:7], h[:, 7:11], h[:, 11], h[:, 12], h[:, 13].astype(np.int32)


def save_kitti_format(sample_id, calib, bbox3d, kitti_output_dir, scores, img_shape, classes="Car"):
    """same arguments as the reference; bbox3d / scores may be CUDA tensors (no per-box host work) or numpy arrays"""
    calib = Calibration(calib)
    dev_boxes = bbox3d if torch.is_tensor(bbox3d) else torch.from_numpy(np.ascontiguousarray(bbox3d, dtype=np.float32)).cuda()
    dev_scores = scores if torch.is_tensor(scores) else torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
    host = detections_to_host(dev_boxes.contiguous(), dev_scores.to(dev_boxes.device), calib.P2, img_shape)
    text = format_kitti_lines(*host, classes=classes)
    path = os.path.join(kitti_output_dir, "%06d.txt" % sample_id)
    with open(path, "w") as f:
        f.write(text)
    return path


def submit_kitti_batch(calibs, img_shapes, boxes3d, scores, select):
    """device half of write_kitti_batch on the CURRENT stream: one launch, one packed (B,M,14) tensor, one asynchronous copy
    into pinned host memory.  Returns (pinned host tensor, event); nothing waits for the GPU here, so several batches can be
    in flight on different streams."""
    C.require_cuda(boxes3d, scores, select)
    B, M = boxes3d.shape[0], boxes3d.shape[1]
    dev = boxes3d.device
    b = boxes3d.contiguous().float()
    p2 = torch.from_numpy(np.stack([Calibration(c).P2.reshape(-1) for c in calibs]).astype(np.float32)).to(dev, non_blocking=True)
    hw = torch.tensor([[float(s[0]), float(s[1])] for s in img_shapes], dtype=torch.float32).to(dev, non_blocking=True)
    img_boxes = torch.empty((B, M, 4), dtype=torch.float32, device=dev)
    alpha = torch.empty((B, M), dtype=torch.float32, device=dev)
    valid = torch.empty((B, M), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        C.check(C.lib().prb_kitti_image_boxes_batch(B, M, C.ptr(b), C.ptr(p2), C.ptr(hw), C.ptr(img_boxes), C.ptr(alpha), C.ptr(valid), C.stream()),
                "kitti_image_boxes_batch")
        keep = valid.float() * select.float()
        packed = torch.cat((b, img_boxes, alpha.unsqueeze(2), scores.float().unsqueeze(2), keep.unsqueeze(2)), dim=2)
        host = torch.empty(packed.shape, dtype=torch.float32, pin_memory=True)
        host.copy_(packed, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        packed.record_stream(torch.cuda.current_stream())
    return host, ev


def collect_kitti_batch(handle, sample_ids=None, kitti_output_dir=None, classes="Car"):
    """host half: wait for the batch's copy, format every scene's text (and write <dir>/<sample_id>.txt when asked)"""
    host, ev = handle
    ev.synchronize()
    h = host.numpy()
    texts = []
    for k in range(h.shape[0]):
        r = h[k]
        texts.append(format_kitti_lines(r[:, 0:7], r[:, 7:11], r[:, 11], r[:, 12], r[:, 13].astype(np.int32), classes))
        if kitti_output_dir is not None:
            with open(os.path.join(kitti_output_dir, "%06d.txt" % int(sample_ids[k])), "w") as f:
                f.write(texts[-1])
    return texts


def write_kitti_batch(sample_ids, calibs, img_shapes, boxes3d, scores, select, kitti_output_dir=None, classes="Car"):
    """A batch of scenes with ONE launch and ONE device-to-host copy: boxes3d (B,M,7), scores (B,M) CUDA tensors,
    select (B,M) bool/0-1 CUDA tensor marking the rows that are detections (e.g. the NMS survivors), rows are written in
    their order.  Returns the list of texts; also writes <dir>/<sample_id>.txt when kitti_output_dir is given."""
    return collect_kitti_batch(submit_kitti_batch(calibs, img_shapes, boxes3d, scores, select), sample_ids, kitti_output_dir, classes)
