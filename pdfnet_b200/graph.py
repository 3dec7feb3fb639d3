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


class PipelinedStep(object):
    """Two-stage software pipeline over CONSECUTIVE batches inside one captured graph.

    The hot path is a chain: the point branch (depth2pcl ... fusion SFT, ``front``) produces ``fuse_feat``
    and the GCN decoder (``back``) consumes nothing else (intaghand_encoder.py:873-874,
    intaghand_decoder.py:180-242).  The two halves stress different resources: ``front`` is a few long
    tensor-core kernels that fill the GPU, ``back`` is ~190 short launches bound by per-launch latency
    that leave most SMs idle.  Replay i therefore runs ``front`` on batch i (current stream) BESIDE
    ``back`` on the hand-over of batch i-1 (second stream) and copies the new hand-over in place once
    both are done.  Every replay still does one ``front`` and one ``back`` worth of work; the decoder
    results of a batch appear one replay after its inputs (``flush()`` drains the last one).

    ``front() -> (handover, front_outputs)``; ``back(handover) -> back_outputs``; ``replay()`` returns
    ``(front_outputs of this batch, back_outputs of the previous batch)``.
    """

    def __init__(self, front, back, warmup=3, back_priority=0):
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        self._back_stream = torch.cuda.Stream(priority=back_priority)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                h, _ = front()
                back(h)
            self.handover = h.clone()                 # the first replay's back stage sees a valid batch
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self._front, self._back = front, back
        self.graph = torch.cuda.CUDAGraph()
        n0 = L.launch_count()
        with torch.cuda.graph(self.graph):
            main = torch.cuda.current_stream()
            self._back_stream.wait_stream(main)       # fork
            with torch.cuda.stream(self._back_stream):
                self.back_outputs = back(self.handover)
            h, self.front_outputs = front()
            main.wait_stream(self._back_stream)       # join: the old hand-over has been consumed
            self.handover.copy_(h)
        self.launches = L.launch_count() - n0
        # the drain step: the back stage alone on the last hand-over
        self._drain = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._drain):
            self.drain_outputs = back(self.handover)

    def replay(self):
        self.graph.replay()
        return self.front_outputs, self.back_outputs

    def flush(self):
        """Back stage of the batch whose front stage ran in the last ``replay()``."""
        self._drain.replay()
        return self.drain_outputs
