"""On-device batch assembly: the host->device boundary of the reference's training / generation loops
(SURVEY.md section 8(f) rank 4).

Reference: `extract_samples(samples, device)` (sbgm/utils.py:405-480) -- called once per batch by
`TrainingPipeline_general.train_batches` (sbgm/training.py:287-292) and the generation loop -- issues one `.to(device).float()`
per entry of the dataset's sample dict and a `torch.cat` of the low-resolution conditions; before that the dataset applied
its transforms sample by sample on the CPU (sbgm/data_modules.py:727-997 with sbgm/special_transforms.py:62-343).

Here the whole dict is staged through ONE pinned host buffer, crosses the bus in ONE copy, and ONE kernel
(`sbgm_assemble_batch`, csrc/batch.cu) writes every float32 output: dtype conversion, the channel concatenation of the LR
conditions and, optionally, each field's forward transform (so raw physical fields can be shipped and normalised on the device).

    from sbgm_danra_b200.batch import extract_samples            # drop-in: same arguments, same 9-tuple
    hr, classifier, lr, lsm_hr, lsm, sdf, topo, hr_pts, lr_pts = extract_samples(samples, device)

    assemble = BatchAssembler(device, transforms={"prcp_hr": PrcpLogTransform(...), "temp_lr": ZScoreTransform(...)})
    hr, classifier, lr, *_ = assemble(raw_samples)               # transforms moved from the dataset to the device

No CPU path: the target must be a CUDA device (as everywhere in this package).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import call

_DT = {torch.float32: 0, torch.float64: 1, torch.float16: 2, torch.bfloat16: 3, torch.int64: 4, torch.int32: 5, torch.int16: 6,
       torch.uint8: 7, torch.bool: 7, torch.int8: 8}


class _Job(ctypes.Structure):            # mirrors sbgm_batch_job (include/sbgm_b200.h)
    _fields_ = [("src", ctypes.c_void_p), ("dst", ctypes.c_void_p), ("count", ctypes.c_longlong), ("inner", ctypes.c_longlong),
                ("dst_stride", ctypes.c_longlong), ("dst_offset", ctypes.c_longlong), ("first_block", ctypes.c_longlong),
                ("dtype", ctypes.c_int), ("log", ctypes.c_int), ("transform", ctypes.c_int), ("pad_", ctypes.c_int),
                ("eps", ctypes.c_float), ("sub", ctypes.c_float), ("mul", ctypes.c_float), ("div", ctypes.c_float),
                ("post_mul", ctypes.c_float), ("post_add", ctypes.c_float)]


def _align(n: int, a: int = 16) -> int:
    return (n + a - 1) // a * a


class _Staging:
    """Two pinned host buffers used alternately (a buffer is rewritten only after the copy that read it has completed) and the
    device buffer the copies land in."""

    def __init__(self, device: torch.device) -> None:
        self.device = device
        self.host: List[Optional[torch.Tensor]] = [None, None]
        self.done: List[Optional[torch.cuda.Event]] = [None, None]
        self.dev: Optional[torch.Tensor] = None
        self.turn = 0

    def acquire(self, nbytes: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
        i = self.turn
        self.turn ^= 1
        if self.done[i] is not None:
            self.done[i].synchronize()
        if self.host[i] is None or self.host[i].numel() < nbytes:
            self.host[i] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        if self.dev is None or self.dev.numel() < nbytes:
            self.dev = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=self.device)
        return self.host[i], self.dev, i

    def release(self, i: int) -> None:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.done[i] = ev


def _run_jobs(device: torch.device, staging: _Staging, items) -> None:
    """items: [(source tensor, destination fp32 tensor, inner, dst_stride, dst_offset, transform or None)]."""
    per = _lib.query("sbgm_batch_chunk_elems")
    # 1. stage the host sources
    offsets, total = [], 0
    for src, *_ in items:
        if src.is_cuda:
            offsets.append(None)
        else:
            offsets.append(total)
            total += _align(src.numel() * src.element_size())
    table_off = total
    total += _align(len(items) * ctypes.sizeof(_Job))
    host, dev, turn = staging.acquire(total)
    keep = []
    for (src, *_), off in zip(items, offsets):
        if off is not None:
            nb = src.numel() * src.element_size()
            flat = src.reshape(-1)
            if flat.dtype == torch.bool:
                flat = flat.view(torch.uint8)
            host[off:off + nb].view(flat.dtype).copy_(flat)
        elif not src.is_contiguous():
            keep.append(src.contiguous())
    # 2. the job table rides in the same copy
    jobs = (_Job * len(items))()
    block = 0
    cont = iter(keep)
    for j, ((src, dst, inner, stride, offset, tf), off) in enumerate(zip(items, offsets)):
        e = jobs[j]
        if off is None:
            s = src if src.is_contiguous() else next(cont)
            e.src = s.data_ptr()
        else:
            e.src = dev.data_ptr() + off
        e.dst, e.count, e.inner, e.dst_stride, e.dst_offset, e.first_block = dst.data_ptr(), src.numel(), inner, stride, offset, block
        e.dtype = _DT[src.dtype]
        if tf is not None:
            lg, eps, sub, mul, div, pm, pa = tf.kernel_params()
            e.log, e.transform, e.eps, e.sub, e.mul, e.div, e.post_mul, e.post_add = int(lg), 1, eps, sub, mul, div, pm, pa
        block += (src.numel() + per - 1) // per
    raw = bytes(jobs)
    host[table_off:table_off + len(raw)].copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    with torch.cuda.device(device):
        dev[:total].copy_(host[:total], non_blocking=True)              # the ONE host->device copy
        staging.release(turn)
        call("sbgm_assemble_batch", dev.data_ptr() + table_off, len(items), block, torch.cuda.current_stream(device).cuda_stream)
    del keep


def _as_cuda(device) -> torch.device:
    if device is None:
        device = torch.device("cuda")
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(f"sbgm_danra_b200.batch assembles batches on CUDA devices only (no CPU path); got {device}")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


class BatchAssembler:
    """`extract_samples` with the copies and conversions of a batch fused (module docstring).  `transforms` maps sample-dict
    keys to forward transforms (special_transforms.Scale / ZScoreTransform / PrcpLogTransform) to apply on the device."""

    def __init__(self, device=None, transforms: Optional[Dict[str, object]] = None) -> None:
        self.device = _as_cuda(device)
        self.transforms = dict(transforms or {})
        self.staging = _Staging(self.device)

    def __call__(self, samples: dict):
        dev, tfs = self.device, self.transforms
        for k, v in samples.items():
            if torch.is_tensor(v) and v.dtype not in _DT and k != "classifier":
                raise TypeError(f"sample '{k}': dtype {v.dtype} is not supported by the batch-assembly kernel")
        # the reference's key rules (utils.py:421-478)
        hr_keys = [k for k in samples.keys() if k.endswith("_hr") and not k.endswith("_original")]
        if "lsm_hr" in hr_keys:
            hr_keys.remove("lsm_hr")
        if len(hr_keys) == 0:
            raise ValueError("No HR image found in samples dictionary.")
        lr_keys = sorted(k for k in samples.keys() if k.endswith("_lr") and not k.endswith("_original"))
        items, out = [], {}

        def whole(key: str, name: str) -> None:
            src = samples.get(key, None)
            if src is None:
                out[name] = None
                return
            dst = torch.empty(src.shape, dtype=torch.float32, device=dev)
            n = max(src.numel(), 1)
            items.append((src, dst, n, n, 0, tfs.get(key)))
            out[name] = dst

        whole(hr_keys[0], "hr")
        if len(lr_keys) == 0:
            out["lr"] = None
        elif len(lr_keys) == 1:
            whole(lr_keys[0], "lr")
        else:                           # torch.cat(dim=1) of the sorted LR conditions: each lands at its channel offset
            shapes = [tuple(samples[k].shape) for k in lr_keys]
            if any(len(s) < 2 or s[0] != shapes[0][0] or s[2:] != shapes[0][2:] for s in shapes):
                raise RuntimeError(f"Sizes of tensors must match except in dimension 1; got {shapes}")
            ctot = sum(s[1] for s in shapes)
            dst = torch.empty((shapes[0][0], ctot) + shapes[0][2:], dtype=torch.float32, device=dev)
            plane = 1
            for d in shapes[0][2:]:
                plane *= d
            coff = 0
            for k, s in zip(lr_keys, shapes):
                items.append((samples[k], dst, s[1] * plane, ctot * plane, coff * plane, tfs.get(k)))
                coff += s[1]
            out["lr"] = dst
        for key in ("lsm_hr", "lsm", "sdf", "topo", "hr_point", "lr_point"):
            whole(key, key)
        items = [it for it in items if it[0].numel() > 0]
        if items:
            _run_jobs(dev, self.staging, items)
        classifier = samples.get("classifier", None)
        if classifier is not None:
            classifier = classifier.to(dev, non_blocking=True)
        return (out["hr"], classifier, out["lr"], out["lsm_hr"], out["lsm"], out["sdf"], out["topo"], out["hr_point"], out["lr_point"])


_DEFAULT: Dict[str, BatchAssembler] = {}


def extract_samples(samples: dict, device=None):
    """Drop-in for `sbgm.utils.extract_samples` (utils.py:405-480): same key rules, same 9-tuple
    (hr_img, classifier, lr_img, lsm_hr, lsm, sdf, topo, hr_points, lr_points), float32 on `device`."""
    dev = _as_cuda(device)
    a = _DEFAULT.get(str(dev))
    if a is None:
        a = _DEFAULT[str(dev)] = BatchAssembler(dev)
    return a(samples)


_TF_STAGING: Dict[str, _Staging] = {}


def apply_transform(sample, transform) -> torch.Tensor:
    """A forward transform on a CUDA tensor: the assembly kernel as a single job."""
    if not isinstance(sample, torch.Tensor) or not sample.is_cuda:
        raise RuntimeError("forward transforms run on CUDA tensors only (no CPU path); move the field to the device, or let "
                           "BatchAssembler apply them while it assembles the batch")
    if sample.dtype not in _DT:
        raise TypeError(f"dtype {sample.dtype} is not supported by the batch-assembly kernel")
    out = torch.empty(sample.shape, dtype=torch.float32, device=sample.device)
    if sample.numel():
        st = _TF_STAGING.setdefault(str(sample.device), _Staging(sample.device))
        n = sample.numel()
        _run_jobs(sample.device, st, [(sample, out, n, n, 0, transform)])
    return out


def install() -> None:
    """Make the reference's loops use this assembler: rebinds `extract_samples` in `sbgm.utils` and in the modules that
    imported it by name (`sbgm.training`, `sbgm.evaluate_sbgm.generation`) if they are importable."""
    import importlib
    import sys
    for name in ("sbgm.utils", "sbgm.training", "sbgm.evaluate_sbgm.generation"):
        mod = sys.modules.get(name)
        if mod is None:
            try:
                mod = importlib.import_module(name)
            except Exception:
                continue
        if hasattr(mod, "extract_samples"):
            mod.extract_samples = extract_samples
