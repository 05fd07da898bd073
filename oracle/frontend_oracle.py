"""TEST INFRASTRUCTURE ONLY (see oracle/svd_oracle.py): torch restatement of the CLIP branch of the conditioning front-end,
StableVideoDiffusionPipeline._encode_image + _resize_with_antialiasing / _gaussian_blur2d / _filter2d
(FTsvd/diffusers-private/diffusers/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:183-229, 643-747).
Pinned against tests/golden/frontend_tiny.npz and frontend_resize.npz (reference-generated).  The product path is
world-in-world_b200/clip.py; only tests import this module."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn.functional as F


def _gaussian(window_size: int, sigma: torch.Tensor) -> torch.Tensor:
    x = (torch.arange(window_size, device=sigma.device, dtype=sigma.dtype) - window_size // 2).expand(sigma.shape[0], -1)
    if window_size % 2 == 0:
        x = x + 0.5
    gauss = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return gauss / gauss.sum(-1, keepdim=True)


def _filter2d(x: torch.Tensor, kernel: torch.Tensor) -> torch.Tensor:
    b, c, h, w = x.shape
    k = kernel[:, None, ...].to(device=x.device, dtype=x.dtype).expand(-1, c, -1, -1)
    kh, kw = k.shape[-2:]
    pad = [(kw - 1) // 2, (kw - 1) - (kw - 1) // 2, (kh - 1) // 2, (kh - 1) - (kh - 1) // 2]
    x = F.pad(x, pad, mode="reflect")
    k = k.reshape(-1, 1, kh, kw)
    x = x.view(-1, k.size(0), x.size(-2), x.size(-1))
    return F.conv2d(x, k, groups=k.size(0), padding=0, stride=1).view(b, c, h, w)


def resize_with_antialiasing(x: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """pipeline :643-667: Gaussian blur with sigma from the scale factor, then bicubic(align_corners=True)."""
    h, w = x.shape[-2:]
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = int(max(2.0 * 2 * sigmas[0], 3)), int(max(2.0 * 2 * sigmas[1], 3))
    ks = (ks[0] + (ks[0] % 2 == 0), ks[1] + (ks[1] % 2 == 0))
    sig = torch.tensor([sigmas], dtype=x.dtype, device=x.device)
    kx, ky = _gaussian(ks[1], sig[:, 1].view(1, 1)), _gaussian(ks[0], sig[:, 0].view(1, 1))
    x = _filter2d(_filter2d(x, kx[..., None, :]), ky[..., None])
    return F.interpolate(x, size=size, mode="bicubic", align_corners=True)


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@torch.no_grad()
def encode_image_clip(image01: torch.Tensor, image_encoder, dtype=None) -> torch.Tensor:
    """pipeline :183-229.  image01 float [B,3,H,W] in [0,1] on the encoder's device -> [B, 1, D] (cond rows only).
    The CLIP input is cast to the encoder's own parameter dtype (pipeline :190, :214)."""
    if dtype is None:
        dtype = next(image_encoder.parameters()).dtype
    x = resize_with_antialiasing(image01 * 2.0 - 1.0, (224, 224))
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN, device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, device=x.device, dtype=x.dtype).view(1, 3, 1, 1)
    x = ((x - mean) / std).to(dtype)
    return image_encoder(x).image_embeds.unsqueeze(1)


