"""Wait-time trace of the pipelined chain kernel (prb_options.mlp_trace): for every chain launch of one backbone forward,
the share of CTA 0's item loop that each role spends in each class of barrier wait.  Usage: python scripts/pipe_trace.py
[key=value prb_options overrides, e.g. mlp_zs=128 mlp_nbuf=1]"""
import ctypes, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench
from pointrcnn_b200 import _cabi as C
from pointrcnn_b200.pointnet2 import pointnet2_modules as pm

opts = {k: (float(v) if "." in v else int(v)) for k, v in (a.split("=") for a in sys.argv[1:])}
dev = torch.device("cuda", 0)
net = bench.build_model(dev)
pc = torch.from_numpy(bench.make_scenes(0, bench.BATCH)).to(dev)
ROLES = [("issuerA", ["x/z_free", "a_full", "b0_full"]), ("issuerB", ["z_free", "ready", "b1_full", "fence", "mma", "commit"]), ("prod0", ["b0_empty"]),
         ("prod1", ["b1_empty"]), ("gather0", ["a_empty"]), ("epi0", ["r_full", "z_full"])]
lib = C.lib()
orig_sa, orig_fp = lib.prb_sa_group_mlp_max_ws, lib.prb_fp_interp_mlp_ws
rows = []


def report(tag, ms):
    buf = (ctypes.c_longlong * 64)()
    lib.prb_debug_pipe_trace(buf)
    t = np.array(buf[:], dtype=np.int64).reshape(8, 8)
    d = {"launch": tag, "ms": round(ms, 4)}
    for i, (name, classes) in enumerate(ROLES):
        tot = max(1, int(t[i, 7]))
        d[name] = {c: round(float(t[i, k]) / tot, 3) for k, c in enumerate(classes)}
        d[name]["kcycles"] = int(tot // 1000)
    rows.append(d)
    print(json.dumps(d))


with torch.no_grad(), C.options(mlp_trace=1, **opts):
    for _ in range(2):
        net(pc)
    torch.cuda.synchronize()
    # time + trace launch by launch: wrap the two module-level launch sites
    from pointrcnn_b200 import prof
    import contextlib

    @contextlib.contextmanager
    def region(name, detail=None):
        if name in ("sa_mlp", "fp_mlp") and detail is not None:
            torch.cuda.synchronize()
            buf = (ctypes.c_longlong * 64)(); lib.prb_debug_pipe_trace(buf)      # clear
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            yield
            b.record(); torch.cuda.synchronize()
            report("%s %s" % (name, detail), a.elapsed_time(b))
        else:
            yield
    prof.region = region
    pm.prof.region = region
    net(pc)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "r2_pipe_trace%s.json" % ("_" + "_".join(sys.argv[1:]) if sys.argv[1:] else "")), "w"), indent=1)
