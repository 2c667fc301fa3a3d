"""Adam / AdamW whose `step()` is ONE kernel launch over every parameter tensor.

The reference builds `torch.optim.Adam(model.parameters(), lr=..., weight_decay=...)` (or AdamW) in
sbgm/training_utils.py:672-698 and calls `self.optimizer.step()` once per batch (sbgm/training.py:407).  torch's default
(foreach) implementation is ~26 launches over the score-UNet's 164 tensors -- 0.5 ms of a 7 ms training step on a B200.
These classes ARE torch.optim.Adam / AdamW (constructor, param_groups, state layout, state_dict / load_state_dict: a
checkpoint's 'optimizer_params' written by either loads into the other); only `step()` is replaced by
`sbgm_adam_step` (csrc/optim.cu), which walks a device table of (param, grad, exp_avg, exp_avg_sq) chunks.

    from sbgm_danra_b200.optim import Adam          # instead of torch.optim.Adam in get_optimizer
    sbgm_danra_b200.optim.install()                 # or: make sbgm.training_utils.get_optimizer build these

No CPU path: parameters must be fp32 CUDA tensors (anything else raises, as everywhere in this package).
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Tuple

import torch

from . import _lib
from ._lib import call


class _Chunk(ctypes.Structure):          # mirrors sbgm_adam_chunk (include/sbgm_b200.h)
    _fields_ = [("param", ctypes.c_void_p), ("grad", ctypes.c_void_p), ("exp_avg", ctypes.c_void_p), ("exp_avg_sq", ctypes.c_void_p),
                ("count", ctypes.c_int), ("pad_", ctypes.c_int)]


class _OneLaunchStep:
    """Mixin: the single-launch step shared by Adam and AdamW (`_decoupled` tells them apart)."""

    _decoupled = False

    def _check_group(self, group: dict) -> None:
        for flag in ("amsgrad", "maximize", "capturable", "differentiable"):
            if group.get(flag):
                raise NotImplementedError(f"sbgm_danra_b200.optim: {flag}=True is not supported by the one-launch step")
        if isinstance(group["lr"], torch.Tensor):
            raise NotImplementedError("sbgm_danra_b200.optim: tensor learning rates are not supported")

    def _table(self, device, rows: List[Tuple[int, int, int, int, int]]) -> torch.Tensor:
        """Device chunk table for `rows` = (param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, numel); cached on the pointers."""
        cache: Dict = self.__dict__.setdefault("_sbgm_tables", {})
        key = (str(device), tuple(rows))
        tab = cache.get(key)
        if tab is None:
            if len(cache) > 8:
                cache.clear()
            per = _lib.query("sbgm_adam_chunk_elems")
            chunks = []
            for p, g, m, v, n in rows:
                for off in range(0, n, per):
                    chunks.append((p + 4 * off, g + 4 * off, m + 4 * off, v + 4 * off, min(per, n - off)))
            arr = (_Chunk * len(chunks))()
            for k, (p, g, m, v, n) in enumerate(chunks):
                arr[k].param, arr[k].grad, arr[k].exp_avg, arr[k].exp_avg_sq, arr[k].count = p, g, m, v, n
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            tab = (host.to(device), len(chunks))
            cache[key] = tab
        return tab

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            self._check_group(group)
            beta1, beta2 = group["betas"]
            by_step: Dict[Tuple[float, str], List[Tuple[int, int, int, int, int]]] = {}
            keep = []
            for p in group["params"]:
                if p.grad is None:
                    continue
                g = p.grad
                if g.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                    raise RuntimeError("sbgm_danra_b200.optim: parameters must be contiguous fp32 CUDA tensors (no CPU path); got "
                                       f"{p.dtype} on {p.device}")
                if g.dtype != torch.float32 or g.device != p.device:
                    raise RuntimeError(f"sbgm_danra_b200.optim: gradient must be fp32 on {p.device}, got {g.dtype} on {g.device}")
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                state = self.state[p]
                if len(state) == 0:      # torch/optim/adam.py _init_group
                    state["step"] = torch.tensor(0.0, dtype=torch.float32)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                state["step"] += 1
                by_step.setdefault((float(state["step"]), str(p.device)), []).append(
                    (p.data_ptr(), g.data_ptr(), state["exp_avg"].data_ptr(), state["exp_avg_sq"].data_ptr(), p.numel()))
            for (step, dev), rows in by_step.items():
                device = torch.device(dev)
                tab, n_chunks = self._table(device, rows)
                with torch.cuda.device(device):
                    call("sbgm_adam_step", tab.data_ptr(), n_chunks, float(group["lr"]), beta1, beta2, group["eps"],
                         group["weight_decay"], int(group.get("decoupled_weight_decay", self._decoupled)), 1.0 - beta1 ** step, 1.0 - beta2 ** step,
                         torch.cuda.current_stream(device).cuda_stream)
            del keep
        return loss


class Adam(_OneLaunchStep, torch.optim.Adam):
    """torch.optim.Adam with a one-launch `step()` (same constructor, state and state_dict)."""
    _decoupled = False


class AdamW(_OneLaunchStep, torch.optim.AdamW):
    """torch.optim.AdamW with a one-launch `step()`."""
    _decoupled = True


def install() -> None:
    """Make the reference's `sbgm.training_utils.get_optimizer` (training_utils.py:672-698, `from torch.optim import Adam,
    AdamW, SGD`) build these classes: rebinds the two names in that module if it is importable."""
    import importlib
    mod = importlib.import_module("sbgm.training_utils")
    mod.Adam, mod.AdamW = Adam, AdamW
