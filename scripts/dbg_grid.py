import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch, synth
from pointrcnn_b200 import _cabi as C
from pointrcnn_b200.pointnet2 import pointnet2_utils as pu
dev = torch.device("cuda:0")
B, N, M = 16, 16384, 4096
xyz = torch.from_numpy(synth.u_kitti(B, N, 5)).to(dev)
idx, known = pu.furthest_point_sample_xyz(xyz, M)
lib = C.lib()
wsb = lib.prb_grid_workspace_bytes(B, M, N)
ws = torch.zeros(wsb, dtype=torch.uint8, device=dev)
d2 = torch.empty((B, N, 3), device=dev); ii = torch.empty((B, N, 3), dtype=torch.int32, device=dev)
C.check(lib.prb_three_nn_grid(B, N, M, C.ptr(xyz), C.ptr(known), C.ptr(d2), C.ptr(ii), None, C.ptr(ws), C.c_size_t(wsb), C.stream()), "nn")
torch.cuda.synchronize()
base = (ws.data_ptr() + 255) // 256 * 256 - ws.data_ptr()
inv_h = ws[base:base + B * 8].view(torch.float64)
print("inv_h", inv_h[:4].tolist(), "h", (1 / inv_h[:2]).tolist())
table = 8192
off = base + ((B * 8 + 255) // 256 * 256) + ((B * 4 + 255) // 256 * 256) + B * table * 4 + B * M * 4
cnt = ws[off:off + 4].view(torch.int32)
print("overflow count", int(cnt[0]), "of", B * N)
print("d3 stats: mean", float(d2[..., 2].mean()), "max", float(d2[..., 2].max()), "frac >= h^2*0.9998:", float((d2[..., 2] >= (1 / inv_h[0]) ** 2 * 0.9998).float().mean()))
