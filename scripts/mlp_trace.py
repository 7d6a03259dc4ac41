"""Phase trace of the chain kernel for one SA level (PRB_MLP_TRACE=1): cycles per phase of CTA 0, tiles 8..31."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
os.environ["PRB_MLP_TRACE"] = "1"
import synth
from pointrcnn_b200 import _cabi as C
from pointrcnn_b200.pointnet2 import pointnet2_modules as pm

dev = torch.device("cuda", 0)
cases = {"SA1": (4096, [0.1, 0.5], [16, 32], [[1, 16, 16, 32], [1, 32, 32, 64]], 1, 16384),
         "SA2": (1024, [0.5, 1.0], [16, 32], [[96, 64, 64, 128], [96, 64, 96, 128]], 96, 4096),
         "SA3": (256, [1.0, 2.0], [16, 32], [[256, 128, 196, 256], [256, 128, 196, 256]], 256, 1024)}
for name, (npoint, radii, ns, mlps, cf, N) in cases.items():
    for sidx in range(2):
        mod = pm.PointnetSAModuleMSG(npoint=npoint, radii=[radii[sidx]], nsamples=[ns[sidx]], mlps=[list(mlps[sidx])], bn=True).to(dev).eval()
        x = torch.from_numpy(synth.u_kitti(16, N, 3)[:, :, :3].copy()).to(dev)
        f = torch.randn(16, cf, N, device=dev)
        with torch.no_grad():
            for _ in range(3):
                mod(x, f)
        buf = (ctypes.c_longlong * 512)()
        C.lib().prb_debug_mlp_trace(buf)
        t = np.array(buf[:], dtype=np.int64).reshape(32, 16)
        L = 3
        k = int((t[:, 2 * L + 1] != 0).sum())          # tiles CTA 0 processed (<= 32)
        lo = min(2, k - 2)
        d = np.diff(t[:k, :2 * L + 2], axis=1)[lo:k - 1]
        nxt = (t[lo + 1:k, 0] - t[lo:k - 1, 2 * L + 1])
        labels = ["gather+A0", "wait D0", "epi0", "wait D1", "epi1", "wait D2", "pool+store"]
        print(name, "scale", sidx, "tiles", k, "cycles/phase:", {l: int(np.median(d[:, i])) for i, l in enumerate(labels)},
              "tile->tile", int(np.median(nxt)), "total/tile", int(np.median(t[lo + 1:k, 0] - t[lo:k - 1, 0])))
