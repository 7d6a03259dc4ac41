"""the per-GPU shapes of the sharded evaluation leg (8 / N scenes per GPU) on ONE GPU: 1, 2 and 4 scenes per step"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import torch
import bench_eval_e2e as E
out = {}
for world in (8, 4, 2):
    r = E.measure(torch.device("cuda:0"), rank=0, world=world, steps=6, warm=2)
    out["%d_scenes_per_gpu" % r["scenes_per_gpu"]] = {"ms_per_step": r["ms_per_step"], "single": r["single_step_in_flight"]["ms_per_step"],
                                                      "scenes_per_s_per_gpu": r["scenes_per_gpu"] / (r["ms_per_step"] * 1e-3)}
    print(world, json.dumps(out["%d_scenes_per_gpu" % r["scenes_per_gpu"]]), flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_eval_shapes.json"), "w"), indent=1)
