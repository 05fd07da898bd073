"""Conditioning front-end of one request: CLIP image embedding + fp32 VAE encoding of the
conditioning frame.  Runs ONCE per request (about 1 % of a rollout's FLOPs) and is the
"next" tier of SURVEY.md section 8(f): for now it calls PyTorch library kernels on the GPU,
exactly as the reference does; the per-step hot loop never touches this module.

Reference: StableVideoDiffusionPipeline._encode_image / _resize_with_antialiasing /
_encode_vae_image (DP/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:183-252,
643-747) and the SD VAE Encoder (DP/models/autoencoders/vae.py:47-182, blocks in
DP/models/unets/unet_2d_blocks.py DownEncoderBlock2D / UNetMidBlock2D).
The CLIP arithmetic lives in third-party `transformers` (parity unpinned at that boundary).
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------- antialiased CLIP resize
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


# ----------------------------------------------------------------------------- SD VAE encoder (fp32)
def _gn(sd: SD, name: str, x: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    return F.group_norm(x, 32, sd[name + ".weight"], sd[name + ".bias"], eps)


def _resnet(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    h = F.conv2d(F.silu(_gn(sd, p + ".norm1", x)), sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], padding=1)
    h = F.conv2d(F.silu(_gn(sd, p + ".norm2", h)), sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], padding=1)
    if (p + ".conv_shortcut.weight") in sd:
        x = F.conv2d(x, sd[p + ".conv_shortcut.weight"], sd[p + ".conv_shortcut.bias"])
    return x + h


def _mid_attention(sd: SD, p: str, x: torch.Tensor) -> torch.Tensor:
    b, c, h, w = x.shape
    t = _gn(sd, p + ".group_norm", x).view(b, c, h * w).transpose(1, 2)
    q = F.linear(t, sd[p + ".to_q.weight"], sd[p + ".to_q.bias"])
    k = F.linear(t, sd[p + ".to_k.weight"], sd[p + ".to_k.bias"])
    v = F.linear(t, sd[p + ".to_v.weight"], sd[p + ".to_v.bias"])
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = F.linear(o, sd[p + ".to_out.0.weight"], sd[p + ".to_out.0.bias"])
    return o.transpose(1, 2).reshape(b, c, h, w) + x


@torch.no_grad()
def vae_encode_mode(sd: SD, image: torch.Tensor, num_blocks: int = 4, layers_per_block: int = 2) -> torch.Tensor:
    """AutoencoderKLTemporalDecoder.encode(x).latent_dist.mode() (autoencoder_kl_temporal_decoder.py:316-341,
    vae.py:140-182): fp32 image [B,3,H,W] in [-1,1] (+ noise aug) -> posterior mean [B,4,H/8,W/8], NOT scaled.
    `sd` holds `encoder.*` and `quant_conv.*` in fp32 on the image's device."""
    x = F.conv2d(image, sd["encoder.conv_in.weight"], sd["encoder.conv_in.bias"], padding=1)
    for i in range(num_blocks):
        for j in range(layers_per_block):
            x = _resnet(sd, f"encoder.down_blocks.{i}.resnets.{j}", x)
        if i < num_blocks - 1:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)                   # downsampling.py: padding == 0 branch
            x = F.conv2d(x, sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"encoder.down_blocks.{i}.downsamplers.0.conv.bias"], stride=2)
    x = _resnet(sd, "encoder.mid_block.resnets.0", x)
    x = _mid_attention(sd, "encoder.mid_block.attentions.0", x)
    x = _resnet(sd, "encoder.mid_block.resnets.1", x)
    x = F.conv2d(F.silu(_gn(sd, "encoder.conv_norm_out", x)), sd["encoder.conv_out.weight"], sd["encoder.conv_out.bias"], padding=1)
    moments = F.conv2d(x, sd["quant_conv.weight"], sd["quant_conv.bias"])
    return moments.chunk(2, dim=1)[0]                                              # DiagonalGaussianDistribution.mode() = mean


def vae_encoder_param_spec(block_out_channels=(128, 256, 512, 512), layers_per_block: int = 2, latent_channels: int = 4):
    """Reference state-dict names / shapes of `encoder.*` + `quant_conv.*` (for seeded random init in the bench)."""
    spec = []
    boc = block_out_channels

    def conv(n, ci, co, k):
        spec.append((n + ".weight", (co, ci, k, k))); spec.append((n + ".bias", (co,)))

    def norm(n, c):
        spec.append((n + ".weight", (c,))); spec.append((n + ".bias", (c,)))

    def res(p, ci, co):
        norm(p + ".norm1", ci); conv(p + ".conv1", ci, co, 3); norm(p + ".norm2", co); conv(p + ".conv2", co, co, 3)
        if ci != co:
            conv(p + ".conv_shortcut", ci, co, 1)

    conv("encoder.conv_in", 3, boc[0], 3)
    out = boc[0]
    for i in range(len(boc)):
        cin, out = out, boc[i]
        for j in range(layers_per_block):
            res(f"encoder.down_blocks.{i}.resnets.{j}", cin if j == 0 else out, out)
        if i < len(boc) - 1:
            conv(f"encoder.down_blocks.{i}.downsamplers.0.conv", out, out, 3)
    res("encoder.mid_block.resnets.0", out, out)
    a = "encoder.mid_block.attentions.0"
    norm(a + ".group_norm", out)
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        spec.append((f"{a}.{nm}.weight", (out, out))); spec.append((f"{a}.{nm}.bias", (out,)))
    res("encoder.mid_block.resnets.1", out, out)
    norm("encoder.conv_norm_out", out)
    conv("encoder.conv_out", out, 2 * latent_channels, 3)
    conv("quant_conv", 2 * latent_channels, 2 * latent_channels, 1)
    return spec
