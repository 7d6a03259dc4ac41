"""Scene-level data parallelism helpers: scenes are independent, so the forward path shards
the batch across ranks with NO data-path collective; training adds one all-reduce over gradients."""
import torch
import torch.distributed as dist


def shard_scenes(num_scenes, rank, world):
    """contiguous [lo, hi) slice of the scene list owned by `rank` (first ranks take the remainder)"""
    base, rem = divmod(num_scenes, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def max_over_ranks(value, device="cuda"):
    """MAX all-reduce of a python float (timings are reported as the max over ranks)"""
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_grads(params, world):
    """one flat all-reduce(sum)/world over the gradients (~12 MB fp32 for the RPN): latency-bound on NVSwitch"""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or world == 1 or not dist.is_initialized():
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        g.copy_(flat[off:off + g.numel()].view_as(g))
        off += g.numel()
