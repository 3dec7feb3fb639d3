"""CUDA-graph capture of a whole hot-path step.

One step is ~30 short kernels; replaying them as ONE graph launch takes the Python / ctypes
launch cost and its jitter out of the step, which matters once several ranks share a host
(SCALE runs).  The C-ABI kernels are enqueued on torch's current stream, never synchronise
and never allocate, so the step is capturable as is; temporaries come from torch's graph
memory pool and stay alive with the graph.
"""
import torch

from . import _lib as L


class CapturedStep(object):
    """``CapturedStep(fn)`` runs ``fn()`` a few times on a side stream (one-time attribute
    calls, weight folding, allocator warm-up), captures one more call, and ``replay()``
    re-executes it on the SAME input buffers (overwrite them in place between replays).
    ``outputs`` are the tensors returned by the captured call; ``launches`` is the number of
    library kernels inside the graph."""

    def __init__(self, fn, warmup=3):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            self.outputs = fn()
        self.launches = L.launch_count() - n0

    def replay(self):
        self.graph.replay()
        return self.outputs
