"""Batch pipelining: consecutive, independent batches on alternating CUDA streams.

Furthest-point sampling is a dependency chain (4095 serial rounds for the first RPN level) that runs as one CTA per
scene (16 of 148 SMs), and nothing else of the SAME batch can run before it ends.  Consecutive batches are
independent (the reference's eval loop processes them one after another), so batch i+1's sampling chain can
overlap batch i's neighbour searches and tensor-core MLPs.  Measured on B200 (RPN backbone, 16 scenes of 16384
points per batch): 4.68 ms per batch back to back on one stream, 2.13 ms with six batches in flight
(profiles/r1_notes.md).  Results are identical to the sequential loop, bit for bit -- only the stream assignment
differs (tests/test_gpu_mlp.py::test_batch_pipeline_matches_sequential).
"""
import torch


class BatchPipeline:
    def __init__(self, fn, inflight=6, device=None, fps_cluster=2):
        """fn: callable(batch_on_device) -> tensor or tuple of tensors; inflight: batches in flight (streams);
        fps_cluster: CTAs per scene of the cluster FPS kernel while the pipeline runs (0 = single-batch heuristic);
        only relevant for scenes the pruned single-CTA kernel does not take (> 16384 points or PRB_FPS_PRUNE=0).
        With several batches in flight SM-time matters more than latency: 2 CTAs per scene take 3.35 ms on 32 SMs,
        4 CTAs 2.53 ms on 64 SMs, the pruned kernel 1.95 ms on 16 SMs."""
        self.fps_cluster = int(fps_cluster)
        self.fn = fn
        self.device = torch.device(device if device is not None else "cuda", torch.cuda.current_device()) \
            if not isinstance(device, torch.device) else device
        self.streams = [torch.cuda.Stream(self.device) for _ in range(max(1, inflight))]
        self._host = {}   # (slot, output index, shape, dtype) -> reusable pinned result buffer (cudaHostAlloc is slow)

    def run(self, batches, to_host=False, keep=True):
        """batches: iterable of tensors (device, or pinned host -> copied inside the pipeline).
        Returns the list of results; with to_host=True results are pinned host tensors (valid after return).
        keep=False drops every result as soon as its work is queued (a consumer inside fn has used it): the
        caching allocator then recycles the output blocks instead of growing by one cudaMalloc per batch."""
        import os
        old_cs = os.environ.get("PRB_FPS_CS")
        if self.fps_cluster > 0 and len(self.streams) > 1:
            os.environ["PRB_FPS_CS"] = str(self.fps_cluster)
        try:
            return self._run(batches, to_host, keep)
        finally:
            if old_cs is None:
                os.environ.pop("PRB_FPS_CS", None)
            else:
                os.environ["PRB_FPS_CS"] = old_cs

    def _run(self, batches, to_host, keep):
        caller = torch.cuda.current_stream(self.device)
        ready = torch.cuda.Event()
        ready.record(caller)
        for s in self.streams:
            s.wait_event(ready)
        results = []
        for i, b in enumerate(batches):
            s = self.streams[i % len(self.streams)]
            with torch.cuda.stream(s):
                x = b if b.is_cuda else b.to(self.device, non_blocking=True)
                out = self.fn(x)
                if to_host:
                    outs = out if isinstance(out, (tuple, list)) else (out,)
                    host = []
                    for j, o in enumerate(outs):
                        key = (i, j, tuple(o.shape), o.dtype)
                        h = self._host.get(key)
                        if h is None:
                            h = self._host[key] = torch.empty(o.shape, dtype=o.dtype, pin_memory=True)
                        h.copy_(o, non_blocking=True)
                        host.append(h)
                    out = host[0] if len(host) == 1 else tuple(host)
                results.append(out if keep else None)
                del out
        for s in self.streams:
            caller.wait_stream(s)
        if to_host:
            for s in self.streams:
                s.synchronize()
        return results
