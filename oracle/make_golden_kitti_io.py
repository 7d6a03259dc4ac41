"""oracle/make_golden_kitti_io.py -- golden vectors for the input pipeline / KITTI output.

Runs the reference's OWN lib/datasets/kitti_rcnn_dataset.py (get_rpn_sample, generate_rpn_training_labels, collate_batch,
imported unchanged from /root/reference) and tools/eval_rcnn.py::save_kitti_format (its function source, compiled as it is)
on synthetic scans, with the dataset's file readers replaced by in-memory scenes (oracle/kitti_io.py::synth_scan), GT-paste
augmentation off (it needs the KITTI GT database), and writes tests/golden/kitti_io.npz.  TEST INFRASTRUCTURE ONLY.
    python oracle/make_golden_kitti_io.py
"""
import ast
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

NPOINTS = 4096
SCENES = [(11, 30000, 6), (12, 4500, 3)]      # (seed, raw points, GT boxes): a subsampled scene and a padded one


def label_line(box, alpha):
    x, y, z, h, w, l, ry = [float(v) for v in box]
    return "Car 0.00 0 %.6f 100.0 100.0 200.0 200.0 %.6f %.6f %.6f %.6f %.6f %.6f %.6f" % (alpha, h, w, l, x, y, z, ry)


def main():
    from pointrcnn_b200 import dropin
    dropin.activate(compat=True)
    import lib.datasets.kitti_rcnn_dataset as K
    from lib.config import cfg
    from lib.utils import calibration, kitti_utils, object3d
    from oracle import kitti_io as KO
    cfg.GT_AUG_ENABLED = False
    cfg.RPN.ENABLED, cfg.RCNN.ENABLED = True, False

    scans = {sid: KO.synth_scan(seed, n, g) for sid, (seed, n, g) in enumerate(SCENES)}
    ds = object.__new__(K.KittiRCNNDataset)
    ds.mode, ds.npoints, ds.random_select = "TRAIN", NPOINTS, True
    ds.classes = ("Background", "Car")
    ds.sample_id_list = [0, 1]
    ds.get_calib = lambda sid: calibration.Calibration(KO.CALIB)
    ds.get_image_shape = lambda sid: KO.IMG_SHAPE
    ds.get_lidar = lambda sid: scans[sid][0]
    ds.get_label = lambda sid: [object3d.Object3d(label_line(b, a)) for b, a in zip(scans[sid][1], scans[sid][2])]

    out = {}
    np.random.seed(2024)
    samples = [ds.get_rpn_sample(0), ds.get_rpn_sample(1)]
    for i, s in enumerate(samples):
        for k in ("pts_input", "pts_rect", "pts_features", "rpn_cls_label", "rpn_reg_label", "gt_boxes3d"):
            out["train%d_%s" % (i, k)] = s[k]
        out["train%d_aug" % i] = np.array(repr(s.get("aug_method", [])))
    batch = ds.collate_batch(samples)
    out["collate_gt_boxes3d"] = batch["gt_boxes3d"]
    out["collate_pts_input_shape"] = np.array(batch["pts_input"].shape)
    out["collate_sample_id"] = np.asarray(batch["sample_id"])
    # intermediate products of scene 0 (for the prepare kernel): rect coordinates and the validity flag
    cal = calibration.Calibration(KO.CALIB)
    rect = cal.lidar_to_rect(scans[0][0][:, 0:3])
    img, depth = cal.rect_to_img(rect)
    out["rect0"], out["valid0"] = rect.astype(np.float32), K.KittiRCNNDataset.get_valid_flag(rect, img, depth, KO.IMG_SHAPE)
    # EVAL mode (labels without augmentation)
    ds.mode = "EVAL"
    np.random.seed(7)
    s = ds.get_rpn_sample(0)
    for k in ("pts_input", "rpn_cls_label", "rpn_reg_label", "gt_boxes3d"):
        out["eval0_%s" % k] = s[k]
    # labels on a dense probe cloud (the static method alone)
    probe_rng = np.random.default_rng(5)
    gt = scans[0][1]
    probe = (gt[probe_rng.integers(0, len(gt), 20000), 0:3] + probe_rng.normal(0, 1.6, (20000, 3))).astype(np.float32)
    probe[:, 1] -= 0.8
    cls, reg = K.KittiRCNNDataset.generate_rpn_training_labels(probe, gt)
    out["probe_pts"], out["probe_cls"], out["probe_reg"] = probe, cls, reg

    # save_kitti_format: the reference's function source, compiled unchanged
    src = open(os.path.join(REF, "tools", "eval_rcnn.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "save_kitti_format"][0]
    ns = dict(np=np, os=os, kitti_utils=kitti_utils, cfg=cfg)
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "eval_rcnn.py", "exec"), ns)
    det_rng = np.random.default_rng(9)
    det = np.concatenate([gt + det_rng.normal(0, 0.05, gt.shape).astype(np.float32),
                          np.array([[0.5, 1.6, 2.5, 1.5, 1.6, 3.9, 0.3], [-2.0, 1.7, 4.0, 1.5, 1.6, 3.9, 1.5]], np.float32)], 0).astype(np.float32)
    scores = det_rng.normal(0, 2, len(det)).astype(np.float32)
    with tempfile.TemporaryDirectory() as d:
        ns["save_kitti_format"](42, cal, det, d, scores, KO.IMG_SHAPE)
        out["kitti_text"] = np.array(open(os.path.join(d, "000042.txt")).read())
    out["kitti_boxes"], out["kitti_scores"] = det, scores

    path = os.path.join(ROOT, "tests", "golden", "kitti_io.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: getattr(v, "shape", None) for k, v in out.items()})
    print(str(out["kitti_text"])[:400])


if __name__ == "__main__":
    main()
