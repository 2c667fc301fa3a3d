"""ctypes binding of the C-ABI library `lib/libsbgm_b200.so` (declared in include/sbgm_b200.h).

There is no CPU fallback: if the library is missing or a kernel launch fails, a RuntimeError is
raised at the call site ("Errors").
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsbgm_b200.so")

FMT_F32, FMT_BF16, FMT_BF16X2 = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU = 0, 1, 2, 3
STEP_COLS = 8
PROJ_STRIDE = 12

_p, _sz, _i, _f, _u64, _u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_float, C.c_uint64, C.c_uint32

# name -> argtypes; every function returns int status unless listed in _RESTYPES
PROTOTYPES = {
    "sbgm_nchw_to_nhwc": [_p, _p, _sz, _i, _i, _i, _i, _i, _p],
    "sbgm_nhwc_to_nchw": [_p, _sz, _i, _p, _i, _i, _i, _i, _p],
    "sbgm_convert": [_p, _sz, _i, _p, _sz, _i, _sz, _p],
    "sbgm_time_embed_project": [_p, _i, _i, _p, _p, _p, _i, _i, _p, _p, _p, _p, _i, _p, _i, _p],
    "sbgm_fourier_embed": [_p, _p, _i, _p, _i, _p],
    "sbgm_cfg_combine": [_p, _p, _f, _p, _sz, _p],
    "sbgm_stem_conv": [_p, _p, _i, _i, _i, _i, _p, _p, _i, _p, _i, _p, _sz, _i, _i, _i, _i, _p],
    "sbgm_conv2d_tc": [_p, _sz, _p, _sz, _p, _p, _sz, _p, _i, _p, _sz, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _i, _p, _p, _sz, _p, _p],
    "sbgm_conv2d_tc_gn_chunks": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
    "sbgm_conv2d_tc_workspace_bytes": [_i, _i, _i, _i, _i, _i, _i, _i, _i, _i],
    "sbgm_conv3x3_c64": [_p, _sz, _p, _sz, _p, _p, _sz, _p, _i, _p, _sz, _i, _i, _i, _i, _i, _p, _i, _p, _p, _i, _p],
    "sbgm_groupnorm_apply": [_p, _sz, _p, _i, _i, _p, _p, _i, _f, _p, _sz, _p, _i, _i, _p, _sz, _i, _i, _i, _i, _p],
    "sbgm_final_gather": [_p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _p],
    "sbgm_conv2d_simt": [_p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p],
    "sbgm_groupnorm": [_p, _sz, _p, _p, _i, _f, _p, _sz, _p, _i, _i, _p, _sz, _i, _i, _i, _i, _p, _p],
    "sbgm_layernorm": [_p, _sz, _p, _p, _f, _p, _sz, _i, _i, _i, _p],
    "sbgm_upsample2x": [_p, _sz, _p, _sz, _i, _i, _i, _i, _i, _p],
    "sbgm_attention": [_p, _sz, _p, _sz, _i, _i, _i, _i, _i, _p],
    "sbgm_final_conv": [_p, _sz, _i, _p, _p, _p, _i, _i, _p, _p, _i, _i, _i, _i, _i, _p],
    "sbgm_philox_normal": [_p, _sz, _u64, _u32, _u64, _p],
    "sbgm_philox_uniform": [_p, _sz, _u64, _u32, _u64, _p],
    "sbgm_sampler_init": [_p, _sz, _f, _u64, _u64, _p],
    "sbgm_sampler_predictor": [_p, _p, _p, _sz, _p, _p, _u32, _u32, _u64, _p],
    "sbgm_sampler_sumsq": [_p, _p, _i, _i, _p],
    "sbgm_sampler_corrector": [_p, _p, _p, _i, _i, _f, _sz, _p, _u32, _u32, _u64, _p],
    "sbgm_select_step_row": [_p, _i, _p, _p, _p],
    "sbgm_dsm_perturb": [_p, _p, _p, _p, _i, _i, _u64, _u32, _u64, _p],
    "sbgm_dsm_loss": [_p, _p, _p, _p, _i, _i, _p, _p, _p],
    "sbgm_groupnorm_scratch_floats": [_i, _i, _i],
    "sbgm_dsm_scratch_floats": [_sz],
    "sbgm_last_error": [],
    "sbgm_version": [],
    "sbgm_device_is_sm100": [],
}
_RESTYPES = {"sbgm_last_error": C.c_char_p, "sbgm_conv2d_tc_workspace_bytes": _sz, "sbgm_groupnorm_scratch_floats": _sz, "sbgm_dsm_scratch_floats": _sz}

_lib: Optional[C.CDLL] = None

# kernel launches issued per entry point (for bench.py's `gpu_launches` claim)
_LAUNCHES = {"sbgm_groupnorm": 2, "sbgm_dsm_loss": 2}   # split-K convolutions add one (not counted: conservative)


class _Stats:
    launches = 0          # kernels launched eagerly or recorded into a CUDA graph through call()


stats = _Stats()


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every symbol of include/sbgm_b200.h (no compute is run)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"sbgm_danra_b200: CUDA library not built ({LIB_PATH}); run `python __graft_entry__.py` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)           # AttributeError if the library does not export it
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; raise RuntimeError with the library's message on failure."""
    lib = load_library()
    status = getattr(lib, name)(*args)
    stats.launches += _LAUNCHES.get(name, 1)
    if status != 0:
        msg = lib.sbgm_last_error()
        raise RuntimeError(f"{name} failed (status {status}): {msg.decode() if msg else '?'}")


def query(name: str, *args):
    return getattr(load_library(), name)(*args)
