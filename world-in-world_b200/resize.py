"""Pillow-exact bicubic resize of uint8 frames on the GPU.

The reference post-processes every generated frame on the CPU with `img.resize(save_size, Image.BICUBIC)`
(downstream/api_models/__init__.py:159) and then quantises twice (to_tensor /255 -> (clip*255).astype(uint8),
:164,229).  Pillow's resampler is integer arithmetic: per output pixel a window [xmin, xmin+n) of normalised bicubic
(a = -0.5) weights, support scaled by the down-sampling factor, converted to 22-bit fixed point, horizontal pass then
vertical pass with an 8-bit intermediate.  `coefficients()` restates that table construction (validated bit-exact
against PIL in tests/test_resize_cpu.py); the two passes run in wiw_resize_bicubic_u8 as integer MACs, so the GPU
result equals PIL's byte for byte.
"""
from __future__ import annotations

from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    return np.where(x < 1.0, ((a + 2.0) * x - (a + 3.0)) * x * x + 1, np.where(x < 2.0, (((x - 5) * x + 8) * x - 4) * a, 0.0))


@lru_cache(maxsize=32)
def coefficients(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(bounds int32 [out, 2] = (first source index, tap count), kk int32 [out, ksize]) of Pillow's precompute_coeffs."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        w = _bicubic((np.arange(xmax) + xmin - center + 0.5) * ss)
        ww = 0.0
        for v in w:                       # Pillow accumulates the normaliser sequentially in double
            ww += float(v)
        if ww != 0.0:
            w = w / ww
        bounds[xx] = (xmin, xmax)
        kk[xx, :xmax] = np.where(w < 0, (w * (1 << PRECISION_BITS) - 0.5).astype(np.int64), (w * (1 << PRECISION_BITS) + 0.5).astype(np.int64))
    return bounds, kk


def requant_lut() -> np.ndarray:
    """uint8 -> uint8 map of the reference's second quantisation: to_tensor (x / 255 in fp32) then
    (np.clip(t, 0, 1) * 255).astype(np.uint8) (truncation) -- api_models/__init__.py:164,229."""
    x = torch.arange(256, dtype=torch.uint8).float().div(255).numpy()
    return (np.clip(x, 0, 1) * 255).astype(np.uint8)


def resize_reference(img: np.ndarray, out_w: int, out_h: int) -> np.ndarray:
    """Pure-numpy evaluation of the two passes (img [H, W, C] uint8) -- the CPU twin of the CUDA kernel, used by tests."""
    H, W, C = img.shape
    bh, kh = coefficients(W, out_w)
    tmp = np.zeros((H, out_w, C), np.uint8)
    for xx in range(out_w):
        x0, n = bh[xx]
        acc = np.full((H, C), 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += img[:, x0 + x, :].astype(np.int64) * int(kh[xx, x])
        tmp[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    bv, kv = coefficients(H, out_h)
    out = np.zeros((out_h, out_w, C), np.uint8)
    for yy in range(out_h):
        y0, n = bv[yy]
        acc = np.full((out_w, C), 1 << (PRECISION_BITS - 1), np.int64)
        for y in range(n):
            acc += tmp[y0 + y].astype(np.int64) * int(kv[yy, y])
        out[yy] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


class GpuResizer:
    """uint8 planes [P, H, W] on the device -> [P, out_h, out_w], Pillow BICUBIC + optional requantisation LUT."""

    def __init__(self, H: int, W: int, out_h: int, out_w: int, device, requant: bool = True):
        self.shape = (H, W, out_h, out_w)
        bh, kh = coefficients(W, out_w)
        bv, kv = coefficients(H, out_h)
        dev = torch.device(device)
        self.bh, self.kh = torch.from_numpy(bh.copy()).to(dev), torch.from_numpy(kh.copy()).to(dev)
        self.bv, self.kv = torch.from_numpy(bv.copy()).to(dev), torch.from_numpy(kv.copy()).to(dev)
        self.lut = torch.from_numpy(requant_lut() if requant else np.arange(256, dtype=np.uint8)).to(dev)
        self.device = dev

    def __call__(self, planes: torch.Tensor) -> torch.Tensor:
        from . import lib as L
        lib = L.require_device()
        H, W, oh, ow = self.shape
        assert planes.dtype == torch.uint8 and planes.device == self.device and planes.is_contiguous() and planes.shape[-2:] == (H, W)
        P = planes.numel() // (H * W)
        tmp = torch.empty(P, H, ow, dtype=torch.uint8, device=self.device)
        out = torch.empty(*planes.shape[:-2], oh, ow, dtype=torch.uint8, device=self.device)
        L.check(lib.wiw_resize_bicubic_u8(planes.data_ptr(), tmp.data_ptr(), out.data_ptr(), P, H, W, oh, ow,
                                          self.bh.data_ptr(), self.kh.data_ptr(), self.kh.shape[1],
                                          self.bv.data_ptr(), self.kv.data_ptr(), self.kv.shape[1],
                                          self.lut.data_ptr(), L.stream_ptr()))
        return out
