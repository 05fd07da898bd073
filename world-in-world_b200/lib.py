"""ctypes binding of libwiw_b200.so (the C ABI declared in include/wiw.h).

The product path has NO fallback: if the shared library is missing or the device is not
sm_100, every op raises.  Tensors are passed as raw device pointers; torch is used only to
own memory and streams.
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# WIW_B200_LIB lets perf scripts A/B an alternative build of the SAME library (development aid, not a fallback)
LIB_PATH = os.environ.get("WIW_B200_LIB", os.path.join(_HERE, "libwiw_b200.so"))
ABI_VERSION = 1


class WiwError(RuntimeError):
    pass


class GemmDesc(ctypes.Structure):
    _fields_ = [
        ("a", ctypes.c_void_p), ("a_dim", ctypes.c_int64 * 5), ("a_stride", ctypes.c_int64 * 5),
        ("w", ctypes.c_void_p), ("k_total", ctypes.c_int64), ("n_total", ctypes.c_int64),
        ("out", ctypes.c_void_p), ("out_dim", ctypes.c_int64 * 4), ("out_stride", ctypes.c_int64 * 4),
        ("res", ctypes.c_void_p), ("res_stride", ctypes.c_int64 * 4),
        ("aux", ctypes.c_void_p), ("aux_stride", ctypes.c_int64 * 4),
        ("bias", ctypes.c_void_p), ("tab", ctypes.c_void_p),
        ("tab_div", ctypes.c_int32), ("tab_mod", ctypes.c_int32),
        ("s_acc", ctypes.c_float), ("s_res", ctypes.c_float), ("s_aux", ctypes.c_float),
        ("bx", ctypes.c_int32), ("by", ctypes.c_int32), ("bn", ctypes.c_int32), ("geglu", ctypes.c_int32),
        ("num_taps", ctypes.c_int32), ("chunks_per_tap", ctypes.c_int32),
        ("tap_cbase", ctypes.c_int32 * 9), ("tap_ox", ctypes.c_int32 * 9),
        ("tap_p", ctypes.c_int32 * 9), ("tap_oy", ctypes.c_int32 * 9),
        ("tab_stride", ctypes.c_int32), ("act", ctypes.c_int32), ("bz", ctypes.c_int32),
    ]


_lib: Optional[ctypes.CDLL] = None

# name -> (restype, argtypes); every symbol include/wiw.h declares
_VP, _I, _I64, _F, _LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_longlong
SYMBOLS = {
    "wiw_last_error": (ctypes.c_char_p, []),
    "wiw_abi_version": (_I, []),
    "wiw_device_ok": (_I, []),
    "wiw_conv_gemm": (_I, [ctypes.POINTER(GemmDesc), _VP]),
    "wiw_groupnorm_workspace_floats": (_I, [_I, _I]),
    "wiw_groupnorm": (_I, [_VP, _VP, _I, _LL, _I, _I, _VP, _VP, _F, _I, _VP, _VP]),
    "wiw_layernorm": (_I, [_VP, _VP, _LL, _I, _VP, _VP, _F, _VP, _I, _I, _VP]),
    "wiw_sampler_prepare": (_I, [_VP, _VP, _VP, _I, _I, _F, _VP]),
    "wiw_sampler_step": (_I, [_VP, _VP, _VP, _I, _I, _I, _F, _F, _F, _F, _VP]),
    "wiw_emb_silu": (_I, [_VP, _VP, _VP, _VP, _I, _I, _VP]),
    "wiw_concat_channels": (_I, [_VP, _VP, _VP, _LL, _I, _I, _VP]),
    "wiw_upsample_nearest2x": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "wiw_attention_temporal": (_I, [_VP, _VP, _I, _I, _I, _I, _VP]),
    "wiw_attention_spatial": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "wiw_softmax_rows": (_I, [_VP, _LL, _I, _F, _VP]),
    "wiw_transpose": (_I, [_VP, _VP, _I, _I, _I, _VP]),
    "wiw_frames_to_u8": (_I, [_VP, _VP, _LL, _LL, _I, _VP]),
    "wiw_vae_finish": (_I, [_VP, _VP, _VP, _VP, _VP, _I, _LL, _I, _VP]),
    "wiw_resize_bicubic_u8": (_I, [_VP, _VP, _VP, _LL, _I, _I, _I, _I, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP]),
    "wiw_resize_separable_f32": (_I, [_VP, _VP, _VP, _LL, _I, _I, _I, _I, _I, _VP, _VP, _I, _VP, _VP, _I, _VP, _VP, _VP]),
    "wiw_patchify": (_I, [_VP, _VP, _I, _I, _I, _I, _I, _VP]),
    "wiw_attention_small": (_I, [_VP, _VP, _I, _I, _I, _I, _F, _VP]),
    "wiw_nchw_to_nhwc": (_I, [_VP, _VP, _LL, _I, _LL, _I, _VP]),
    "wiw_nhwc_to_nchw": (_I, [_VP, _VP, _LL, _I, _LL, _I, _VP]),
}


# launch accounting for bench.py's `gpu_launches` (kernels launched by each C-ABI call)
COUNTERS = {"gpu_launches": 0}
KERNELS_PER_CALL = {"wiw_groupnorm": 3, "wiw_resize_bicubic_u8": 2, "wiw_resize_separable_f32": 2}
_NO_LAUNCH = {"wiw_last_error", "wiw_abi_version", "wiw_device_ok", "wiw_groupnorm_workspace_floats"}


def load() -> ctypes.CDLL:
    """dlopen the library (works on a CPU-only box: it does not link libcuda)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WiwError(f"{LIB_PATH} not found -- run `python __graft_entry__.py build` (no CPU fallback exists)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)          # AttributeError if a declared symbol is not exported
            fn.restype = res
            fn.argtypes = args
            if name in _NO_LAUNCH:
                continue

            def counted(*a, _fn=fn, _k=KERNELS_PER_CALL.get(name, 1)):
                COUNTERS["gpu_launches"] += _k
                return _fn(*a)
            setattr(lib, name, counted)
        if lib.wiw_abi_version() != ABI_VERSION:
            raise WiwError(f"ABI mismatch: library {lib.wiw_abi_version()} vs binding {ABI_VERSION}")
        _lib = lib
    return _lib


def require_device() -> ctypes.CDLL:
    lib = load()
    if not torch.cuda.is_available() or not lib.wiw_device_ok():
        raise WiwError("wiw_b200 needs an sm_100 (B200) CUDA device; there is no CPU or library fallback")
    return lib


def check(status: int) -> None:
    if status != 0:
        raise WiwError(f"libwiw_b200 status {status}: {load().wiw_last_error().decode()}")


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())
