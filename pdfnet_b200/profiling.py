"""Optional per-stage CUDA-event timing (used by bench.py for the roofline lines).

Disabled by default: ``stage(name)`` is then a no-op context manager.  When enabled,
each stage records a start/stop event pair on torch's current stream (the stream the
C-ABI kernels are enqueued on); ``summary()`` synchronises once and returns
{name: (calls, median_ms)}.
"""
import contextlib

import torch

_enabled = False
_events = []


def enable(flag=True):
    global _enabled
    _enabled = flag
    _events.clear()


@contextlib.contextmanager
def stage(name):
    if not _enabled:
        yield
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    try:
        yield
    finally:
        b.record()
        _events.append((name, a, b))


def summary():
    """{stage: (calls, median_ms)} — the median is robust against one-off allocator / first-touch stalls."""
    torch.cuda.synchronize()
    per = {}
    for name, a, b in _events:
        per.setdefault(name, []).append(a.elapsed_time(b))
    _events.clear()
    out = {}
    for name, ts in per.items():
        ts.sort()
        out[name] = (len(ts), ts[len(ts) // 2])
    return out
