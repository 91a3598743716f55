"""Fork / join of a side CUDA stream next to the caller's current stream.

The hot loops are chains of small kernels whose neighbours are often independent (actor vs. critic MLP, fused MDP step vs.
taxel synthesis, dgrad vs. wgrad GEMM).  Issuing the independent chain on a side stream lets the GPU overlap them -- eagerly
and, because event record / wait are capturable, as parallel branches of a CUDA graph."""
from __future__ import annotations

import contextlib
import gc
import os

import torch


@contextlib.contextmanager
def graph_capture(graph: "torch.cuda.CUDAGraph", **kwargs):
    """``torch.cuda.graph`` with Python's cyclic garbage collector held off for the duration of the capture.

    Destroying an old ``CUDAGraph`` (cudaGraphExecDestroy, release of its memory pool) is not permitted while a stream of the
    thread is capturing and INVALIDATES the capture in progress; an unreachable engine of an earlier run that the cyclic collector
    happens to free in the middle of a capture does exactly that (torch >= 2.9 no longer collects on entry).  So: collect once
    before the capture starts, then keep the automatic collector off until it ends."""
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, **kwargs):
            yield
    finally:
        if was_enabled:
            gc.enable()


class SideStream:
    def __init__(self, device):
        self.device = torch.device(device)
        enabled = os.environ.get("LT_SIDE_STREAMS", "1") != "0"  # LT_SIDE_STREAMS=0: everything on the caller's stream
        self.stream = torch.cuda.Stream(device=self.device) if (self.device.type == "cuda" and enabled) else None

    def mark(self):
        """A fork point on the current stream for a LATER ``forked(after=mark)``: the body then depends only on what was enqueued up to
        here, although it is enqueued (launched, or created as a graph node) after whatever the current stream received in between."""
        if self.stream is None:
            return None
        ev = torch.cuda.Event()
        ev.record()
        return ev

    @contextlib.contextmanager
    def forked(self, after=None):
        """``with side.forked(): ...`` enqueues the body on the side stream, ordered after everything already enqueued on the
        current stream (``after``: only after that earlier ``mark()``).  Call ``join()`` before the current stream consumes what the
        body produced."""
        if self.stream is None:
            yield
            return
        ev = after
        if ev is None:
            ev = torch.cuda.Event()
            ev.record()
        self.stream.wait_event(ev)
        with torch.cuda.stream(self.stream):
            yield

    def join(self):
        if self.stream is None:
            return
        ev = torch.cuda.Event()
        ev.record(self.stream)
        torch.cuda.current_stream().wait_event(ev)
