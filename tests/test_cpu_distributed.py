"""CPU suite: the N>1 scene-sharding logic of bench.py under gloo, world_size 2 (no data-path collective:
scenes are independent; only the timing reduction and the gradient all-reduce helper talk)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch, torch.distributed as dist
from pointrcnn_b200.parallel_utils import shard_scenes, max_over_ranks, allreduce_grads
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
lo, hi = shard_scenes(16, rank, world)
assert hi - lo == 8 and lo == rank * 8
lo2, hi2 = shard_scenes(5, rank, world)
assert (lo2, hi2) == ((0, 3) if rank == 0 else (3, 5))
t = max_over_ranks(float(rank + 1), device='cpu')
assert t == 2.0
lin = torch.nn.Linear(4, 2)
for p in lin.parameters():
    p.grad = torch.full_like(p, float(rank + 1))
allreduce_grads(lin.parameters(), world)
assert all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in lin.parameters())
# bucketed reducer: two ranks, different data, same weights -> averaged gradients equal the full-batch gradient
from pointrcnn_b200.parallel_utils import GradBucketReducer
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3))
ref = torch.nn.Sequential(torch.nn.Linear(6, 32), torch.nn.ReLU(), torch.nn.Linear(32, 32), torch.nn.ReLU(), torch.nn.Linear(32, 3))
ref.load_state_dict(net.state_dict())
red = GradBucketReducer(net.parameters(), world=world, bucket_mb=0.002)      # several buckets
assert len(red.buckets) >= 3
g = torch.Generator().manual_seed(5)
x = torch.randn(8, 6, generator=g)
for step in range(2):                                # twice: reset() must keep the gradients inside the buckets
    red.reset()
    net(x[rank * 4:(rank + 1) * 4]).pow(2).mean().backward()
    red.finish()
ref(x).pow(2).mean().backward()
for a, b in zip(net.parameters(), ref.parameters()):
    assert torch.allclose(a.grad, b.grad, rtol=1e-5, atol=1e-6), (a.grad - b.grad).abs().max()
    assert a.grad.data_ptr() >= red.buckets[red._owner[a]]["flat"].data_ptr()
dist.barrier()
if rank == 0: print('DIST_OK')
""" % ROOT


def test_gloo_world2_sharding_and_reductions(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=240)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
