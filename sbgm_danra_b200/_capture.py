"""CUDA-graph capture with the two precautions every capture site of this package needs."""
from __future__ import annotations

import contextlib
import gc

import torch


@contextlib.contextmanager
def graph_capture(graph: "torch.cuda.CUDAGraph", **kwargs):
    """`torch.cuda.graph(graph, **kwargs)` with

    * `capture_error_mode="thread_local"`: CUDA calls of OTHER threads (a DataLoader's pin-memory thread: cudaHostAlloc, event
      queries) do not invalidate the capture;
    * a full garbage-collection pass before the capture and the cyclic collector paused during it.  A dead reference cycle that
      owns CUDA graphs (an earlier model with its cached sampler plan or training runner) is otherwise finalised whenever the
      collector happens to run -- and destroying a graph while a stream is capturing is "not permitted" and invalidates the
      capture.  (torch.cuda.graph used to collect at entry; since torch 2.x it does so only under
      `torch.compiler.config.force_cudagraph_gc`.  Seen as an order-dependent failure of one GPU test late in a long process.)
    """
    gc.collect()
    was_enabled = gc.isenabled()
    gc.disable()
    try:
        with torch.cuda.graph(graph, capture_error_mode="thread_local", **kwargs):
            yield
    finally:
        if was_enabled:
            gc.enable()
