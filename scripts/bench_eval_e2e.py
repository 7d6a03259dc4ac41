"""BASELINE configs[4]: end-to-end two-stage evaluation (tools/eval_rcnn.py eval_one_epoch_joint, :459-640) on synthetic
KITTI-shaped scans, global batch 8 sharded over the ranks (8 / world scenes per GPU):

    raw scans in pinned host memory -> H2D -> input pipeline (calibration, validity, 16384-point draw; TEST mode)
    -> RPN (backbone, heads, proposal layer: 100 RoIs per scene) -> roipool3d + RCNN stage -> decode, score threshold,
    rotated NMS -> image boxes / alpha -> D2H of the detections -> KITTI result text per scene (host formatter)

measure() returns one JSON-able dict; run as a script for a single-GPU record.  No data-path collective: scenes are independent."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GLOBAL_BATCH, RAW_POINTS, NPOINTS = 8, 60000, 16384
P2 = np.array([[721.5377, 0.0, 609.5593, 44.85728], [0.0, 721.5377, 172.854, 0.2163791], [0.0, 0.0, 1.0, 0.002745884]], np.float32)
CALIB = dict(P2=P2,
             R0=np.array([[0.9999239, 0.00983776, -0.007445048], [-0.009869795, 0.9999421, -0.004278459],
                          [0.007402527, 0.004351614, 0.9999631]], np.float32),
             Tr_velo2cam=np.array([[0.007533745, -0.9999714, -0.000616602, -0.004069766], [0.01480249, 0.0007280733, -0.9998902, -0.07631618],
                                   [0.9998621, 0.00752379, 0.01480755, -0.2717806]], np.float32))
IMG_SHAPE = (375, 1242, 3)


def synth_raw_scan(seed, n=RAW_POINTS):
    """raw lidar-frame scan whose camera frustum part fills PC_AREA_SCOPE roughly uniformly (~2/3 of the points are valid)"""
    rng = np.random.default_rng(seed)
    z = rng.uniform(0.5, 72.0, n)
    x = rng.uniform(-1.0, 1.0, n) * np.minimum(0.9 * z, 42.0)
    y = rng.uniform(-1.3, 3.3, n)
    return np.stack([z + 0.27, -x, -y - 0.08, rng.random(n)], 1).astype(np.float32)


def build(dev):
    from pointrcnn_b200.point_rcnn import PointRCNNInference
    torch.manual_seed(21)
    model = PointRCNNInference(input_channels=1).to(dev).eval()
    g = torch.Generator().manual_seed(22)
    for m in model.modules():
        if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
            m.running_mean.copy_((torch.randn(m.running_mean.shape, generator=g) * 0.1).to(dev))
            m.running_var.copy_((torch.rand(m.running_var.shape, generator=g) + 0.5).to(dev))
    with torch.no_grad():
        model.rpn.rpn_reg_layer[-1].conv.weight.mul_(0.2)      # usable box sizes on random weights
    return model


def measure(dev, rank=0, world=1, steps=10, warm=3, barrier=lambda: None, max_over_ranks=lambda v, device=None: v, inflight=3):
    from pointrcnn_b200 import _cabi as C
    from pointrcnn_b200.datasets import kitti_output
    from pointrcnn_b200.datasets.kitti_rcnn_dataset import RPNInputPipeline
    per_rank = max(1, GLOBAL_BATCH // world)
    model = build(dev)
    pipe = RPNInputPipeline(npoints=NPOINTS, mode="TEST", draw="device", device=dev)
    pool = [[dict(lidar=torch.from_numpy(synth_raw_scan(9000 + 100 * p + rank * per_rank + i)).pin_memory(), calib=CALIB, img_shape=IMG_SHAPE) for i in range(per_rank)]
            for p in range(4)]
    h2d = sum(s["lidar"].numel() * 4 for s in pool[0])
    stats = {"detections": 0, "text_bytes": 0, "d2h": 0}

    def submit(i):
        scans = pool[i % len(pool)]
        batch = pipe.prepare_batch(scans, seed=i)
        out = model(batch["pts_input"])
        boxes, raw, select = model.detections_device(out)          # no host round trip up to here
        return kitti_output.submit_kitti_batch([CALIB] * len(scans), [IMG_SHAPE] * len(scans), boxes, raw, select)

    def collect(h):
        texts = kitti_output.collect_kitti_batch(h)
        stats["detections"] += sum(t.count("\n") for t in texts); stats["text_bytes"] += sum(len(t) for t in texts)
        stats["d2h"] += h[0].numel() * 4

    streams = [torch.cuda.Stream(dev) for _ in range(max(1, inflight))]

    def run(first, count, depth):
        """`depth` steps in flight: step i runs on stream i % depth, its text is formatted while later steps compute"""
        pending = []
        for i in range(first, first + count):
            with torch.cuda.stream(streams[i % depth]):
                pending.append(submit(i))
            if len(pending) >= depth:
                collect(pending.pop(0))
        while pending:
            collect(pending.pop(0))

    res = {}
    with torch.no_grad():
        run(0, warm, 1)
        run(warm, 2 * max(1, inflight), max(1, inflight))
        torch.cuda.synchronize()
        for name, depth in (("single", 1), ("pipelined", max(1, inflight))):
            reps = []
            for rep in range(3):                                       # the region is host-paced: median of three repeats
                stats.update(detections=0, text_bytes=0, d2h=0)
                lc0 = C.launch_count()
                barrier(); torch.cuda.synchronize()
                t0 = time.perf_counter()
                run(100 + rep * steps, steps, depth)
                torch.cuda.synchronize()
                wall = (time.perf_counter() - t0) * 1e3 / steps        # the last result is on the host: wall clock == step time
                barrier()
                reps.append(max_over_ranks(wall, device=dev))
                launches = (C.launch_count() - lc0) // steps
            ms = sorted(reps)[1]
            res[name] = {"ms_per_step": ms, "value": per_rank * world / (ms * 1e-3), "steps_in_flight": depth, "ms_per_step_min_max": [min(reps), max(reps)]}
    ms, wall = res["pipelined"]["ms_per_step"], res["pipelined"]["ms_per_step"]
    return {"what": "two-stage evaluation end to end (BASELINE configs[4]): raw scans (pinned host) -> input pipeline -> RPN -> RCNN -> "
                    "rotated NMS -> KITTI result text; global batch %d, %d scene(s) per GPU" % (GLOBAL_BATCH, per_rank),
            "ms_per_step": ms, "value": per_rank * world / (ms * 1e-3), "unit": "scenes/s", "single_step_in_flight": res["single"],
            "steps_in_flight": res["pipelined"]["steps_in_flight"],
            "timing": "host wall clock around K steps incl. the last result's D2H and text (sync on both sides), max over ranks",
            "scenes_per_gpu": per_rank, "steps": steps, "raw_points_per_scan": RAW_POINTS,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": stats["d2h"] // steps,
            "detections_per_scene": stats["detections"] / (steps * per_rank), "text_bytes_per_scene": stats["text_bytes"] / (steps * per_rank),
            "gpu_launches_per_step": launches, "collective": None}


if __name__ == "__main__":
    r = measure(torch.device("cuda:0"))
    print(json.dumps(r))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(r, open(os.path.join(ROOT, "gpurun_out", "r2_eval_e2e.json"), "w"), indent=1)
