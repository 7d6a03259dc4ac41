"""Per-kernel-family device timing with CUDA events on the launching stream (used by bench.py).

    prof.enable(); ...run...; torch.cuda.synchronize(); times = prof.collect()   # {family: [ms, launches]}

Disabled (the default) it costs one attribute check per native call."""
import contextlib

import torch

from . import config

_on = False
_events = []


def enable():
    global _on
    _on = True
    _events.clear()


def disable():
    global _on
    _on = False


@contextlib.contextmanager
def region(name, detail=None):
    """detail: optional shape string; with config prof_detail it is appended to the family name"""
    if _on and detail is not None and config.get("prof_detail"):
        with region("%s %s" % (name, detail)):   # the detailed line nests inside the family line
            with region(name):
                yield
        return
    if not _on:
        yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _events.append((name, a, b))


def collect():
    """call after a synchronize; returns {name: [total_ms, count]} and clears the log"""
    out = {}
    for name, a, b in _events:
        t = out.setdefault(name, [0.0, 0])
        t[0] += a.elapsed_time(b)
        t[1] += 1
    _events.clear()
    return out
