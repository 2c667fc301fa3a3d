"""Data-parallel DSM training: one process per GPU, replicated weights, NCCL gradient all-reduce.

The reference trains on one device (sbgm/training.py:246-422); BASELINE.json's C4 shards the minibatch over
the GPUs of one box.  `TrainEngine` writes every parameter gradient as a view of ONE flat fp32 buffer
(state-dict order, time-projection parameters last), so the exchange is an in-place bucketed
`all_reduce` of that buffer -- no flatten / unflatten copies:

  * buckets are contiguous ranges of the flat buffer, which is laid out in FORWARD order (train_engine.flat_order: time
    projections, stem, encoder stages, decoder blocks, final layer); backward produces gradients from the END of the buffer
    down, so buckets complete from the high end and each is launched on a communication stream as soon as every gradient in it
    has been enqueued -- overlapping the remaining backward kernels.  Only the lowest bucket (stem + time projections, a few
    hundred KB) is exposed after the last backward kernel;
  * gradients are averaged like torch DDP (`ReduceOp.AVG` inside NCCL; sum then scale on backends without it), so
    `loss.backward(); optimizer.step()` in the reference training loop needs no change.  `grad_dtype="bf16"` exchanges a
    bfloat16 copy of each bucket (half the NVLink bytes; the local fp32 buffer receives the rounded average);
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


def _produced_last(name: str) -> bool:
    """Parameters whose gradients come out of the final kernels of TrainEngine.backward (one launch set for all of them)."""
    return "time_projection_layer" in name or name.endswith("label_emb.weight")


class GradBucketer:
    """Tracks which contiguous buckets of a flat gradient buffer are complete.

    `layout`: [(name, offset, numel)] sorted by offset.  `expected`: names that will be touched in a step."""

    def __init__(self, layout: Sequence[Tuple[str, int, int]], total: int, bucket_elems: int, expected: Optional[Sequence[str]] = None):
        self.layout, self.total = list(layout), total
        self.bounds: List[Tuple[int, int]] = []
        # Walk from the END of the buffer (filled first by backward) in full-sized buckets; the START of the buffer holds what
        # backward produces last (stem, time projections), so the lowest buckets are made SMALL (1/32 and 1/8 of a bucket): the
        # only exchange that cannot hide behind remaining backward kernels is then ~1 MB instead of whatever the walk left over.
        # The very first bucket is exactly the group that the LAST kernels of backward produce in one go (time projections and
        # label embedding: train_engine.flat_order puts them first): mixed into a size-cut bucket they held back the stem /
        # layer-1 gradients that were ready 50 us earlier, and two exchanges (4.2 MB + 1.0 MB) ran after backward instead of one.
        tail_sizes = [max(1, bucket_elems // 32), max(1, bucket_elems // 8)]
        head_edges, acc, k = [0], 0, 0
        late_end = 0
        for name, off, numel in self.layout:
            if _produced_last(name):
                late_end = min(total, off + (numel + 63) // 64 * 64)
            else:
                break
        if late_end > 0:
            head_edges.append(late_end)
        for name, off, numel in self.layout:
            if k >= len(tail_sizes):
                break
            if off < late_end:
                continue
            acc += numel
            if acc >= tail_sizes[k]:
                head_edges.append(off + (numel + 63) // 64 * 64 if off + (numel + 63) // 64 * 64 <= total else total)
                acc, k = 0, k + 1
        lo_limit = head_edges[-1]
        edges = [total]
        acc = 0
        for name, off, numel in reversed(self.layout):
            if off < lo_limit:
                break
            acc += numel
            if acc >= bucket_elems:
                edges.append(off)
                acc = 0
        edges = sorted(set(edges) | set(e for e in head_edges if e <= total) | {0})
        self.bounds = [(edges[i], edges[i + 1]) for i in range(len(edges) - 1) if edges[i + 1] > edges[i]]
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
        self.late: List[str] = []      # touched outside the expected set after their bucket had already been exchanged
        self.unexpected: List[str] = []   # every gradient touched outside the expected set (the set grows by these in finish())

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
            elif self.expected is not None:
                self.unexpected.append(name)
                if self.launched[self.bucket_of[name]]:
                    self.late.append(name)     # its bucket has gone already: exchanged on its own in finish()
        return ready

    def remaining(self) -> List[int]:
        return [b for b in range(len(self.bounds)) if not self.launched[b]]


class GradSync:
    """The hook `TrainEngine.backward` drives: `begin(engine)`, `progress(names)` after every tape step,
    `finish()` once all gradients are enqueued.  All-reduces on a side stream, averages, then joins."""

    def __init__(self, group=None, bucket_bytes: int = BUCKET_BYTES, sync_bn: bool = False, grad_dtype: str = "fp32") -> None:
        if grad_dtype not in ("fp32", "bf16"):
            raise ValueError(f"grad_dtype must be 'fp32' or 'bf16', got {grad_dtype!r}")
        self.group, self.bucket_elems, self.sync_bn, self.grad_dtype = group, max(1, bucket_bytes // 4), sync_bn, grad_dtype
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.native_avg = dist.is_initialized() and dist.get_backend(group) == "nccl"
        self.expected: Optional[List[str]] = None
        self.bucketer: Optional[GradBucketer] = None
        self.stream: Optional[torch.cuda.Stream] = None
        self.flat: Optional[torch.Tensor] = None
        self.works: list = []
        self.stats = {"buckets": 0, "overlapped": 0}
        self.last_plan: Optional[dict] = None    # the exchange of the most recent step (what a captured graph replays)
        self._step_overlapped: List[int] = []

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
        self._step_overlapped = []

    def _average(self, view: torch.Tensor) -> None:
        """In-place average of `view` over the group."""
        buf = view.to(torch.bfloat16) if self.grad_dtype == "bf16" else view
        if self.native_avg:
            dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)      # the division happens inside the NCCL kernel
        else:
            dist.all_reduce(buf, group=self.group)
            buf.mul_(1.0 / self.world)
        if buf is not view:
            view.copy_(buf)

    def _launch_range(self, lo: int, hi: int) -> None:
        view = self.flat[lo:hi]
        if self.flat.is_cuda:
            self.stream.wait_stream(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self.stream):
                self._average(view)
        else:
            self._average(view)

    def _launch(self, b: int, overlapped: bool) -> None:
        self.stats["buckets"] += 1
        self.stats["overlapped"] += int(overlapped)
        self._launch_range(*self.bucketer.bounds[b])

    def progress(self, names: Sequence[str]) -> None:
        if self.world == 1:
            return
        for b in self.bucketer.touch(names):
            self._step_overlapped.append(b)
            self._launch(b, True)

    def finish(self) -> None:
        if self.world == 1:
            return
        tail = self.bucketer.remaining()
        for b in tail:
            self.bucketer.launched[b] = True
            self._launch(b, False)
        mb = lambda b: round((self.bucketer.bounds[b][1] - self.bucketer.bounds[b][0]) * 4 / 2 ** 20, 3)
        self.last_plan = {"buckets": len(self.bucketer.bounds), "overlapped": len(self._step_overlapped),
                          "overlapped_mb": [mb(b) for b in self._step_overlapped], "after_backward_mb": [mb(b) for b in tail],
                          "late_gradients": len(self.bucketer.late), "dtype": self.grad_dtype,
                          "op": "ncclAllReduce AVG" if self.native_avg else "all_reduce SUM then scale"}
        offs = {name: (off, numel) for name, off, numel in self.bucketer.layout}
        for name in self.bucketer.late:
            off, numel = offs[name]
            self._launch_range(off, off + numel)
        if self.bucketer.unexpected and self.expected is not None:
            self.expected = sorted(set(self.expected) | set(self.bucketer.unexpected))
            self.bucketer = None            # rebuilt with the enlarged set by the next begin()
        if self.flat.is_cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.stream)
        if self.expected is None and self.bucketer is not None:
            self.expected = sorted(self.bucketer.seen)
        self.flat = None


def _drop_captured_steps(model) -> None:
    """Captured training graphs have the exchange (or its absence) baked in: forget them so the next step re-captures."""
    runners = getattr(model, "__dict__", {}).get("_train_runners")
    if runners:
        runners.clear()


def attach(model, group=None, bucket_bytes: int = BUCKET_BYTES, sync_bn: bool = False, grad_dtype: str = "fp32") -> GradSync:
    """Make `loss.backward()` through `model` (a ScoreNet of this package) average gradients over `group`.

    sync_bn=True additionally synchronises the BatchNorm batch statistics (forward: all-gather of the per-chunk partial
    sums; backward: all-gather of the per-sample gradient sums), so that N ranks with B/N samples each reproduce the
    reference's single-process batch-B statistics exactly; the default is DDP-style rank-local statistics."""
    sync = GradSync(group, bucket_bytes, sync_bn, grad_dtype)
    model._grad_sync = sync
    _drop_captured_steps(model)
    return sync


def detach(model) -> None:
    model._grad_sync = None
    _drop_captured_steps(model)


def broadcast_parameters(model, src: int = 0, group=None) -> None:
    """Replicate rank `src`'s parameters and buffers (done once before training, like DDP's constructor)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(model.parameters()) + list(model.buffers()):
            # t.detach() shares t's version counter (t.data does not): the in-place write invalidates the packed-weight
            # caches (score_unet._EngineCache keys on data_ptr + _version) of engines built before the broadcast
            dist.broadcast(t.detach(), src=src, group=group)
    _drop_captured_steps(model)
