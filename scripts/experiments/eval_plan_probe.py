"""evaluation leg with the backbone's side-stream plan (config.enable_plan) on / off"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import bench_eval_e2e as E
from pointrcnn_b200 import config
out = {}
for plan in (False, True):
    with config.override(enable_plan=plan):
        r = E.measure(torch.device("cuda:0"), steps=10, warm=3)
    out["plan_%d" % int(plan)] = {"pipelined_ms": r["ms_per_step"], "single_ms": r["single_step_in_flight"]["ms_per_step"]}
    print(plan, json.dumps(out["plan_%d" % int(plan)]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_eval_plan_probe.json"), "w"), indent=1)
