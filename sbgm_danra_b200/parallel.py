"""Data-parallel DSM training: one process per GPU, replicated weights, NCCL gradient all-reduce.

The reference trains on one device (sbgm/training.py:246-422); BASELINE.json's C4 shards the minibatch over
the GPUs of one box.  `TrainEngine` writes every parameter gradient as a view of ONE flat fp32 buffer
(state-dict order, time-projection parameters last), so the exchange is an in-place bucketed
`all_reduce` of that buffer -- no flatten / unflatten copies:

  * buckets are contiguous ranges of the flat buffer; backward finishes the decoder (the END of the buffer)
    first, so buckets complete from the high end down and each is launched on a communication stream as soon
    as every gradient in it has been enqueued -- overlapping the remaining backward kernels;
  * gradients are averaged (sum / world) like torch DDP, so `loss.backward(); optimizer.step()` in the reference
    training loop needs no change;
  * the first step all-reduces the whole buffer at the end and records which parameters receive gradients
    (the final block's unused time projection etc. never do); later steps use that set.

Usage:  parallel.attach(model, group=None, sync_bn=False)   then train as usual.  Ranks must use distinct batches; the Philox
stream is keyed by global member index via score_sampling.set_ensemble_shard(first_member=rank * B_local).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

BUCKET_BYTES = 32 << 20


class GradBucketer:
    """Tracks which contiguous buckets of a flat gradient buffer are complete.

    `layout`: [(name, offset, numel)] sorted by offset.  `expected`: names that will be touched in a step."""

    def __init__(self, layout: Sequence[Tuple[str, int, int]], total: int, bucket_elems: int, expected: Optional[Sequence[str]] = None):
        self.layout, self.total = list(layout), total
        self.bounds: List[Tuple[int, int]] = []
        start, n = 0, 0
        # walk from the END of the buffer (filled first by backward) so that early buckets are full-sized
        edges = [total]
        acc = 0
        for name, off, numel in reversed(self.layout):
            acc += numel
            if acc >= bucket_elems:
                edges.append(off)
                acc = 0
        if edges[-1] != 0:
            edges.append(0)
        edges = sorted(set(edges))
        self.bounds = [(edges[i], edges[i + 1]) for i in range(len(edges) - 1)]
        self.bucket_of: Dict[str, int] = {}
        for name, off, numel in self.layout:
            for b, (lo, hi) in enumerate(self.bounds):
                if lo <= off < hi:
                    self.bucket_of[name] = b
                    break
        self.expected = None if expected is None else set(expected)
        self.reset()

    def reset(self) -> None:
        self.pending = [0] * len(self.bounds)
        if self.expected is not None:
            for name in self.expected:
                self.pending[self.bucket_of[name]] += 1
        self.seen = set()
        self.launched = [False] * len(self.bounds)

    def touch(self, names: Sequence[str]) -> List[int]:
        """Mark gradients as enqueued; returns the buckets that just became complete (expected set known)."""
        ready = []
        for name in names:
            if name in self.seen:
                continue
            self.seen.add(name)
            if self.expected is not None and name in self.expected:
                b = self.bucket_of[name]
                self.pending[b] -= 1
                if self.pending[b] == 0 and not self.launched[b]:
                    self.launched[b] = True
                    ready.append(b)
        return ready

    def remaining(self) -> List[int]:
        return [b for b in range(len(self.bounds)) if not self.launched[b]]


class GradSync:
    """The hook `TrainEngine.backward` drives: `begin(engine)`, `progress(names)` after every tape step,
    `finish()` once all gradients are enqueued.  All-reduces on a side stream, averages, then joins."""

    def __init__(self, group=None, bucket_bytes: int = BUCKET_BYTES, sync_bn: bool = False) -> None:
        self.group, self.bucket_elems, self.sync_bn = group, max(1, bucket_bytes // 4), sync_bn
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.expected: Optional[List[str]] = None
        self.bucketer: Optional[GradBucketer] = None
        self.stream: Optional[torch.cuda.Stream] = None
        self.flat: Optional[torch.Tensor] = None
        self.works: list = []
        self.stats = {"buckets": 0, "overlapped": 0}

    def begin(self, flat: torch.Tensor, layout: Sequence[Tuple[str, int, int]]) -> None:
        self.flat = flat
        if self.world == 1:
            return
        if flat.is_cuda:
            if self.stream is None:
                self.stream = torch.cuda.Stream(device=flat.device)
            if not torch.cuda.is_current_stream_capturing():
                flat.record_stream(self.stream)
        if self.bucketer is None or self.bucketer.total != flat.numel() or (self.bucketer.expected is None and self.expected is not None):
            self.bucketer = GradBucketer(layout, flat.numel(), self.bucket_elems, self.expected)
        self.bucketer.reset()
        self.works = []

    def _launch(self, b: int, overlapped: bool) -> None:
        lo, hi = self.bucketer.bounds[b]
        view = self.flat[lo:hi]
        self.stats["buckets"] += 1
        self.stats["overlapped"] += int(overlapped)
        if self.flat.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.stream):
                dist.all_reduce(view, group=self.group)
                view.mul_(1.0 / self.world)
        else:
            dist.all_reduce(view, group=self.group)
            view.mul_(1.0 / self.world)

    def progress(self, names: Sequence[str]) -> None:
        if self.world == 1:
            return
        for b in self.bucketer.touch(names):
            self._launch(b, True)

    def finish(self) -> None:
        if self.world == 1:
            return
        for b in self.bucketer.remaining():
            self.bucketer.launched[b] = True
            self._launch(b, False)
        if self.flat.is_cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        if self.expected is None:
            self.expected = sorted(self.bucketer.seen)
        self.flat = None


def attach(model, group=None, bucket_bytes: int = BUCKET_BYTES, sync_bn: bool = False) -> GradSync:
    """Make `loss.backward()` through `model` (a ScoreNet of this package) average gradients over `group`.

    sync_bn=True additionally synchronises the BatchNorm batch statistics (forward: all-gather of the per-chunk partial
    sums; backward: all-gather of the per-sample gradient sums), so that N ranks with B/N samples each reproduce the
    reference's single-process batch-B statistics exactly; the default is DDP-style rank-local statistics."""
    sync = GradSync(group, bucket_bytes, sync_bn)
    model._grad_sync = sync
    return sync


def detach(model) -> None:
    model._grad_sync = None


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    """Replicate rank `src`'s parameters and buffers (done once before training, like DDP's constructor)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=group)
