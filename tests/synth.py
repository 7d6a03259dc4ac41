"""Seeded synthetic inputs shared by tests, golden generation and bench.py."""
import numpy as np


def u_kitti(B, N, seed, channels=3):
    """uniform in PC_AREA_SCOPE [-40,40]x[-1,3]x[0,70.4] (tools/cfgs/default.yaml:18) (+ intensity-0.5)"""
    rng = np.random.default_rng(seed)
    lo = np.array([-40.0, -1.0, 0.0], dtype=np.float32)
    hi = np.array([40.0, 3.0, 70.4], dtype=np.float32)
    pts = (lo + (hi - lo) * rng.random((B, N, 3), dtype=np.float32)).astype(np.float32)
    if channels > 3:
        inten = rng.random((B, N, channels - 3), dtype=np.float32) - np.float32(0.5)
        pts = np.concatenate([pts, inten], axis=2)
    return np.ascontiguousarray(pts)


def u_cube(B, N, seed):
    """uniform in the unit cube: dense balls (early exit + full neighbourhoods)"""
    return np.random.default_rng(seed).random((B, N, 3), dtype=np.float32)


def dup_cloud(B, N, seed, unique=None):
    """tie stress: only `unique` distinct points, tiled and shuffled (RoI-style duplicate padding)"""
    unique = unique or max(1, N // 4)
    rng = np.random.default_rng(seed)
    base = u_kitti(B, unique, seed + 7)
    out = np.empty((B, N, 3), dtype=np.float32)
    for b in range(B):
        out[b] = base[b][rng.integers(0, unique, size=N)]
    return out


def boxes3d(N, seed, centres=48):
    """(N,7) [x,y,z,h,w,l,ry] clustered proposals around `centres` objects; scores (N) unique"""
    rng = np.random.default_rng(seed)
    cx = rng.uniform(-35, 35, centres); cz = rng.uniform(5, 65, centres); cy = rng.uniform(1.2, 2.0, centres)
    base_ry = rng.uniform(-np.pi, np.pi, centres)
    which = rng.integers(0, centres, N)
    b = np.empty((N, 7), dtype=np.float32)
    b[:, 0] = cx[which] + rng.normal(0, 0.3, N)
    b[:, 1] = cy[which] + rng.normal(0, 0.1, N)
    b[:, 2] = cz[which] + rng.normal(0, 0.3, N)
    size = np.array([1.526, 1.629, 3.883])  # h, w, l  (default.yaml:19 CLS_MEAN_SIZE)
    b[:, 3:6] = size[None, :] * (1 + rng.uniform(-0.15, 0.15, (N, 3)))
    b[:, 6] = base_ry[which] + rng.normal(0, 0.1, N)
    scores = (rng.random(N) + np.arange(N) * 1e-7).astype(np.float32)
    return b, scores


def to_bev(b):
    out = np.empty((b.shape[0], 5), dtype=np.float32)
    hl, hw = b[:, 5] / np.float32(2), b[:, 4] / np.float32(2)
    out[:, 0], out[:, 1], out[:, 2], out[:, 3], out[:, 4] = b[:, 0] - hl, b[:, 2] - hw, b[:, 0] + hl, b[:, 2] + hw, b[:, 6]
    return out


def sorted_bev(N, seed):
    b, s = boxes3d(N, seed)
    order = np.argsort(-s, kind="stable")
    return np.ascontiguousarray(to_bev(b)[order])
