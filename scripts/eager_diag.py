"""why is the eager single-stream forward slow in some bench runs?  CPU vs GPU time of 10 eager forwards: plain, with the
prof regions, with per-shape detail; cProfile of the slowest variant"""
import os, sys, time, cProfile, pstats, io
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from pointrcnn_b200 import prof, config
dev = torch.device("cuda:0")
net = bench.build_model(dev)
pool = [torch.from_numpy(bench.make_scenes(i, bench.BATCH)).to(dev) for i in range(4)]
flush = torch.empty(64 * 1024 * 1024, device=dev)


def run(tag, n=10, with_prof=False, detail=False, profile=False):
    ctx = config.override(prof_detail=detail)
    with torch.no_grad(), ctx:
        for i in range(2):
            net(pool[i % 4])
        torch.cuda.synchronize()
        if with_prof:
            prof.enable()
        pr = cProfile.Profile() if profile else None
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
        t0 = time.perf_counter()
        if pr: pr.enable()
        for i, (a, b) in enumerate(ev):
            flush.fill_(1.0)
            a.record(); net(pool[i % 4]); b.record()
        if pr: pr.disable()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if with_prof:
            prof.disable(); fam = prof.collect()
        gpu = sum(a.elapsed_time(b) for a, b in ev) / n
    print("%-28s cpu launch %.2f ms/fwd, wall incl. drain %.2f ms/fwd, gpu %.2f ms/fwd" % (tag, (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3, gpu), flush=True)
    if with_prof:
        print("     families:", {k: round(v[0] / n, 3) for k, v in fam.items() if " " not in k})
    if pr:
        s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18); print(s.getvalue()[:3500])


run("plain")
run("prof regions", with_prof=True)
run("prof regions + detail", with_prof=True, detail=True)
run("plain again")
run("prof detail, cProfile", with_prof=True, detail=True, profile=True)
