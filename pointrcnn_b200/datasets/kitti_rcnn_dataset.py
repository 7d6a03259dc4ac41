"""RPN input pipeline on the device: mirror of lib/datasets/kitti_rcnn_dataset.py:246-394 (get_rpn_sample,
generate_rpn_training_labels), :198-219 (get_valid_flag), :513-570 (data_augmentation, stage 1) and :1104-1137
(collate_batch) -- SURVEY.md 8(f) rank 4.

The reference prepares ONE scene per `__getitem__` in numpy (calibration products, a boolean-mask compaction, np.random
draws, two Delaunay triangulations per GT box for the labels) and stacks scenes in `collate_batch`.  Here a whole batch
of raw scans goes through four launches (csrc/kitti_io.cu) and comes out as the tensors the network and the loss read:

    pipe = RPNInputPipeline(npoints=16384, mode="TRAIN")
    batch = pipe.prepare_batch(scans, seed=step)          # dict of CUDA tensors, same keys as collate_batch

File IO (KITTI .bin / calib / label readers, the GT-paste database) stays with the caller: a scan is
`dict(lidar=(n,4) float32 numpy array or (pinned) torch tensor, calib=Calibration|dict, img_shape=(H,W[,3]),
gt_boxes3d=(g,7), gt_alpha=(g,))`.

Two ways to draw the npoints sample:
  draw="device" (default)  counter-based hashes on the device, no host round trip; same distribution as the reference
  draw="numpy"             the reference's own np.random calls in the reference's order (needs the valid flags on the
                           host): bit-for-bit the reference's sample for a given np.random state -- used by the parity tests
There is no CPU path: the kernels are required (RuntimeError without the library / a CUDA device).
"""
import ctypes

import numpy as np
import torch

from .. import _cabi as C

PC_AREA_SCOPE = ((-40.0, 40.0), (-1.0, 3.0), (0.0, 70.4))      # tools/cfgs/default.yaml:18


class Calibration(object):
    """lib/utils/calibration.py:24-41 for an in-memory dict {'P2','R0','Tr_velo2cam'} (fp32 like get_calib_from_file)"""

    def __init__(self, calib):
        if isinstance(calib, Calibration):
            calib = dict(P2=calib.P2, R0=calib.R0, Tr_velo2cam=calib.V2C)
        self.P2 = np.asarray(calib["P2"], dtype=np.float32).reshape(3, 4)
        self.R0 = np.asarray(calib["R0"], dtype=np.float32).reshape(3, 3)
        self.V2C = np.asarray(calib["Tr_velo2cam"], dtype=np.float32).reshape(3, 4)

    def lidar_to_rect_matrix(self):
        """(4,3) fp32: rect = [x y z 1] . M  (calibration.py:57: np.dot(self.V2C.T, self.R0.T))"""
        return np.dot(self.V2C.T, self.R0.T).astype(np.float32)

    def pack(self, img_shape, scope=None):
        row = np.zeros(32, dtype=np.float32)
        row[0:12] = self.lidar_to_rect_matrix().reshape(-1)
        row[12:24] = self.P2.reshape(-1)
        row[24], row[25] = float(img_shape[0]), float(img_shape[1])
        if scope is not None:
            row[26:32] = np.asarray(scope, dtype=np.float32).reshape(-1)
        return row


def _dev(device):
    device = torch.device(device if device is not None else "cuda")
    if device.type != "cuda":
        raise RuntimeError("pointrcnn_b200: the input pipeline runs on a CUDA device (there is no CPU path)")
    return device


def _i32(t):
    return t if t.dtype == torch.int32 else t.to(torch.int32)


def generate_rpn_training_labels(pts_rect, gt_boxes3d, extra_width=0.2, gt_count=None):
    """kitti_rcnn_dataset.py:355-391.  numpy (N,3),(G,7) -> numpy (N,) int32, (N,7) float32 like the reference's static
    method; CUDA tensors (B,N,3),(B,G,7) -> CUDA tensors (B,N) int32, (B,N,7) (batched, zero rows of a padded batch skipped
    unless gt_count (B,) says how many rows are real)."""
    as_numpy = isinstance(pts_rect, np.ndarray)
    if as_numpy:
        dev = _dev(None)
        p = torch.from_numpy(np.ascontiguousarray(pts_rect, dtype=np.float32)).to(dev).unsqueeze(0)
        g = torch.from_numpy(np.ascontiguousarray(gt_boxes3d, dtype=np.float32).reshape(-1, 7)).to(dev).unsqueeze(0)
        gt_count = torch.tensor([g.size(1)], dtype=torch.int32, device=dev)
    else:
        p, g = pts_rect.contiguous(), gt_boxes3d.contiguous()
        C.require_cuda(p, g)
        if p.dtype != torch.float32 or g.dtype != torch.float32:
            raise RuntimeError("generate_rpn_training_labels: float32 tensors expected")
    B, N, _ = p.shape
    G = g.size(1)
    cls = torch.empty((B, N), dtype=torch.int32, device=p.device)
    reg = torch.empty((B, N, 7), dtype=torch.float32, device=p.device)
    if gt_count is not None:
        gt_count = _i32(gt_count.to(p.device)).contiguous()
    with torch.cuda.device(p.device):
        C.check(C.lib().prb_rpn_training_labels(B, N, G, C.ptr(p), C.ptr(g) if G else None, C.ptr(gt_count),
                                                ctypes.c_float(extra_width), C.ptr(cls), C.ptr(reg), C.stream()), "rpn_training_labels")
    if as_numpy:
        return cls[0].cpu().numpy(), reg[0].cpu().numpy()
    return cls, reg


def draw_augmentation(rng=np.random, aug_list=("rotation", "scaling", "flip"),
                      aug_prob=(0.5, 0.5, 0.5), rot_range=18, mustaug=False):
    """the random draws of data_augmentation (kitti_rcnn_dataset.py:520-568) in the reference's order:
    -> (angle or None, scale or None, flip bool, aug_method list)"""
    aug_enable = 1 - rng.rand(3)
    if mustaug:
        aug_enable[0] = aug_enable[1] = -1
    angle = scale = None
    flip = False
    method = []
    if "rotation" in aug_list and aug_enable[0] < aug_prob[0]:
        angle = rng.uniform(-np.pi / rot_range, np.pi / rot_range)
        method.append(["rotation", angle])
    if "scaling" in aug_list and aug_enable[1] < aug_prob[1]:
        scale = rng.uniform(0.95, 1.05)
        method.append(["scaling", scale])
    if "flip" in aug_list and aug_enable[2] < aug_prob[2]:
        flip = True
        method.append("flip")
    return angle, scale, flip, method


def augment_gt_boxes3d(gt_boxes3d, gt_alpha, angle, scale, flip):
    """the GT-box half of data_augmentation (stage 1): a handful of boxes per scene, host numpy with the reference's
    float64 intermediates (kitti_rcnn_dataset.py:528-560, kitti_utils.py:32-42)"""
    b = np.array(gt_boxes3d, dtype=np.float32, copy=True)
    if angle is not None and b.shape[0]:
        c, s = np.cos(angle), np.sin(angle)
        x, z = b[:, 0].astype(np.float64), b[:, 2].astype(np.float64)
        b[:, 0], b[:, 2] = x * c - z * s, x * s + z * c
        beta = np.arctan2(b[:, 2], b[:, 0])
        b[:, 6] = np.sign(beta) * np.pi / 2 + gt_alpha - beta
    if scale is not None:
        b[:, 0:6] = b[:, 0:6] * scale
    if flip:
        b[:, 0] = -b[:, 0]
        b[:, 6] = np.sign(b[:, 6]) * np.pi - b[:, 6]
    return b


def draw_choice_numpy(pts_depth, npoints, rng=np.random):
    """the npoints draw of kitti_rcnn_dataset.py:285-303 with the reference's np.random calls in its order;
    pts_depth = rect z of the VALID points; returns indices into the valid points"""
    n = len(pts_depth)
    if npoints < n:
        near = pts_depth < 40.0
        far_idxs = np.where(near == 0)[0]
        near_idxs = np.where(near == 1)[0]
        picked = rng.choice(near_idxs, npoints - len(far_idxs), replace=False)
        choice = np.concatenate((picked, far_idxs), axis=0) if len(far_idxs) > 0 else picked
        rng.shuffle(choice)
    else:
        choice = np.arange(0, n, dtype=np.int32)
        if npoints > n:
            extra = rng.choice(choice, npoints - n, replace=False)
            choice = np.concatenate((choice, extra), axis=0)
        rng.shuffle(choice)
    return choice


def collate_batch(batch):
    """kitti_rcnn_dataset.py:1104-1137 (RPN mode): gt_boxes3d zero-padded to the longest list, arrays stacked,
    ints / floats to arrays, everything else listed"""
    out = {}
    bs = len(batch)
    for key in batch[0].keys():
        if key == "gt_boxes3d":
            mg = max(len(batch[k][key]) for k in range(bs))
            g = np.zeros((bs, mg, 7), dtype=np.float32)
            for i in range(bs):
                g[i, :len(batch[i][key]), :] = batch[i][key]
            out[key] = g
        elif isinstance(batch[0][key], np.ndarray):
            out[key] = np.stack([batch[k][key] for k in range(bs)], axis=0)
        else:
            vals = [batch[k][key] for k in range(bs)]
            if isinstance(batch[0][key], int):
                vals = np.array(vals, dtype=np.int32)
            elif isinstance(batch[0][key], float):
                vals = np.array(vals, dtype=np.float32)
            out[key] = vals
    return out


class RPNInputPipeline(object):
    def __init__(self, npoints=16384, mode="TRAIN", use_intensity=True, reduce_by_range=True, pc_area_scope=PC_AREA_SCOPE,
                 aug_data=True, aug_method_list=("rotation", "scaling", "flip"), aug_method_prob=(0.5, 0.5, 0.5), aug_rot_range=18,
                 random_select=True, draw="device", device=None):
        assert mode in ("TRAIN", "EVAL", "TEST") and draw in ("device", "numpy")
        self.npoints, self.mode, self.use_intensity = int(npoints), mode, bool(use_intensity)
        self.reduce_by_range, self.scope = bool(reduce_by_range), pc_area_scope
        self.aug_data = bool(aug_data) and mode == "TRAIN"
        self.aug = (tuple(aug_method_list), tuple(aug_method_prob), aug_rot_range)
        self.random_select, self.draw = bool(random_select), draw
        self.device = _dev(device)

    # ---------------------------------------------------------------- device stages
    def _upload(self, scans):
        B = len(scans)
        counts = [int(s["lidar"].shape[0]) for s in scans]
        offsets = np.zeros(B + 1, dtype=np.int32)
        offsets[1:] = np.cumsum(counts)
        total = int(offsets[-1])
        stride = int(scans[0]["lidar"].shape[1])
        dev = self.device
        lidar = torch.empty((max(total, 1), stride), dtype=torch.float32, device=dev)
        for s, o in zip(scans, offsets[:-1]):      # one DMA per scan, straight from the caller's (ideally pinned) memory
            src = s["lidar"]
            n = int(src.shape[0])
            if n:
                src = src if torch.is_tensor(src) else torch.from_numpy(np.ascontiguousarray(src, dtype=np.float32))
                lidar[o:o + n].copy_(src, non_blocking=True)
        calib = np.stack([Calibration(s["calib"]).pack(s["img_shape"], self.scope if self.reduce_by_range else None) for s in scans])
        return (B, total, stride, torch.from_numpy(offsets).to(dev, non_blocking=True), lidar,
                torch.from_numpy(calib).to(dev, non_blocking=True), offsets)

    def prepare_batch(self, scans, seed=0, rng=np.random):
        """-> dict of CUDA tensors with collate_batch's keys: pts_input (B,npoints,3|4), pts_rect (B,npoints,3),
        pts_features (B,npoints,1), and unless mode == 'TEST' gt_boxes3d (B,max_g,7) zero padded, rpn_cls_label (B,npoints) int32,
        rpn_reg_label (B,npoints,7); plus `choice` (B,npoints) int32 raw-point indices and `aug_method` per scene."""
        lib = C.lib()
        dev = self.device
        B, total, stride, offsets, lidar, calib, offsets_h = self._upload(scans)
        npoints = self.npoints
        with torch.cuda.device(dev):
            rect = torch.empty((max(total, 1), 3), dtype=torch.float32, device=dev)
            flags = torch.empty(max(total, 1), dtype=torch.uint8, device=dev)
            counts = torch.empty((B, 2), dtype=torch.int32, device=dev)
            C.check(lib.prb_kitti_prepare_points(B, total, C.ptr(offsets), C.ptr(lidar), stride, C.ptr(calib), int(self.reduce_by_range),
                                                 C.ptr(rect), C.ptr(flags), C.ptr(counts), C.stream()), "kitti_prepare_points")
            choice = torch.empty((B, npoints), dtype=torch.int32, device=dev)
            status = torch.empty(B, dtype=torch.int32, device=dev)
            if self.draw == "device":
                cand = torch.empty(max(total, 1), dtype=torch.int32, device=dev)
                C.check(lib.prb_kitti_draw_points(B, total, C.ptr(offsets), C.ptr(flags), npoints, ctypes.c_uint(int(seed) & 0xffffffff),
                                                  C.ptr(cand), C.ptr(choice), C.ptr(status), C.stream()), "kitti_draw_points")
            else:
                # the reference's np.random stream: flags and depths come back to the host (one sync per batch)
                fl = flags.cpu().numpy()
                depth = rect[:, 2].cpu().numpy()
                rows = []
                for b in range(B):
                    f = fl[offsets_h[b]:offsets_h[b + 1]]
                    valid = np.nonzero(f & 1)[0]
                    if self.mode == "TRAIN" or self.random_select:
                        ch = draw_choice_numpy(depth[offsets_h[b]:offsets_h[b + 1]][valid], npoints, rng)
                    else:
                        raise RuntimeError("draw='numpy' without random_select returns ragged scenes; use the reference loader")
                    rows.append(valid[ch].astype(np.int32))
                choice.copy_(torch.from_numpy(np.stack(rows)))
                status.zero_()
            # augmentation parameters: the reference draws them per scene after the sample (host, a few numbers)
            aug_rows, aug_methods, gts = None, [], []
            if self.aug_data:
                aug_rows = np.zeros((B, 4), dtype=np.float64)
            for b, s in enumerate(scans):
                g = np.asarray(s.get("gt_boxes3d", np.zeros((0, 7), np.float32)), dtype=np.float32).reshape(-1, 7)
                if self.aug_data:
                    angle, scale, flip, method = draw_augmentation(rng=rng, aug_list=self.aug[0], aug_prob=self.aug[1], rot_range=self.aug[2])
                    aug_rows[b] = (np.cos(angle) if angle is not None else 1.0, np.sin(angle) if angle is not None else 0.0,
                                   float(np.float32(scale)) if scale is not None else 1.0, 1.0 if flip else 0.0)
                    alpha = np.asarray(s.get("gt_alpha", np.zeros(len(g), np.float32)), dtype=np.float32)
                    g = augment_gt_boxes3d(g, alpha, angle, scale, flip)
                    aug_methods.append(method)
                gts.append(g)
            aug_t = torch.from_numpy(aug_rows).to(dev, non_blocking=True) if aug_rows is not None else None
            channels = 4 if (self.use_intensity and stride > 3) else 3
            pts_input = torch.empty((B, npoints, channels), dtype=torch.float32, device=dev)
            pts_rect = torch.empty((B, npoints, 3), dtype=torch.float32, device=dev)
            inten = torch.empty((B, npoints), dtype=torch.float32, device=dev)
            C.check(lib.prb_kitti_gather_points(B, npoints, C.ptr(offsets), C.ptr(rect), C.ptr(lidar), stride, C.ptr(choice), C.ptr(aug_t),
                                                C.ptr(status), channels, C.ptr(pts_input), C.ptr(pts_rect), C.ptr(inten), C.stream()), "kitti_gather_points")
            out = dict(pts_input=pts_input, pts_rect=pts_rect, pts_features=inten.unsqueeze(-1), choice=choice, status=status,
                       valid_counts=counts)
            if self.mode != "TEST":
                mg = max([len(g) for g in gts] + [0])
                gpad = np.zeros((B, mg, 7), dtype=np.float32)
                for b, g in enumerate(gts):
                    gpad[b, :len(g)] = g
                gt_t = torch.from_numpy(gpad).to(dev, non_blocking=True)
                gcnt = torch.tensor([len(g) for g in gts], dtype=torch.int32).to(dev, non_blocking=True)
                out["gt_boxes3d"] = gt_t
                # like get_rpn_sample, every mode but TEST carries the per-point labels (kitti_rcnn_dataset.py:343-352)
                out["rpn_cls_label"], out["rpn_reg_label"] = generate_rpn_training_labels(pts_rect, gt_t, gt_count=gcnt)
            if self.aug_data:
                out["aug_method"] = aug_methods
        return out
