"""RPN training step (BASELINE configs[2]) under library settings: cuDNN autotuning, channels-last weights; per-family device time"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from pointrcnn_b200 import prof
from pointrcnn_b200.train.step import RPNTrainer, synthetic_labels

dev = torch.device("cuda:0")
pool = [torch.from_numpy(bench.make_scenes(16 * i, bench.BATCH)).to(dev) for i in range(2)]
labels = [synthetic_labels(p, seed=i) for i, p in enumerate(pool)]
out = {}
for name, bm, cl in (("default", False, False), ("cudnn_benchmark", True, False), ("cudnn_benchmark+channels_last", True, True)):
    torch.backends.cudnn.benchmark = bm
    tr = RPNTrainer(input_channels=bench.CHANNELS - 3, device=dev, world=1)
    if cl:
        tr.model.to(memory_format=torch.channels_last)
    try:
        for i in range(3):
            tr.step(pool[i % 2], *labels[i % 2], grad_norm_clip=1.0)
    except Exception as e:
        import traceback
        out[name] = {"failed": traceback.format_exc()[-400:]}
        print(name, out[name], flush=True)
        continue
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for i in range(6):
        tr.step(pool[i % 2], *labels[i % 2], grad_norm_clip=1.0)
    b.record(); torch.cuda.synchronize()
    out[name] = {"ms_per_step": a.elapsed_time(b) / 6}
    if name == "default":
        # forward / backward split and the forward's families
        a, m, b = (torch.cuda.Event(True) for _ in range(3))
        prof.enable()
        a.record()
        loss, _ = tr.forward_loss(pool[0], *labels[0])
        m.record()
        loss.backward()
        b.record(); torch.cuda.synchronize()
        prof.disable()
        out[name].update(forward_ms=a.elapsed_time(m), backward_ms=m.elapsed_time(b),
                         forward_families_ms={k: v[0] for k, v in prof.collect().items()})
    print(name, json.dumps(out[name]), flush=True)
    tr.reducer.remove(); del tr
    torch.cuda.empty_cache()
print(json.dumps(out))
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r2_train_probe.json"), "w"), indent=1)
