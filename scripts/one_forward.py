"""one backbone forward of the bench workload after warm-ups (target of the ncu captures)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
dev = torch.device("cuda:0")
net = bench.build_model(dev)
pc = torch.from_numpy(bench.make_scenes(0, bench.BATCH)).to(dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
with torch.no_grad():
    for _ in range(n - 1):          # warm-ups (the first one also measures the chain plans)
        net(pc)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()     # ncu --profile-from-start off: only the last forward is captured
    net(pc)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("done")
