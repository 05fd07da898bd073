"""fp32 VAE encoding of the conditioning frame (once per request, 2.6 TFLOP).

The reference deliberately runs this stage in fp32 (`force_upcast`: the SD VAE overflows fp16 on real checkpoints), so the
default path keeps fp32 arithmetic on PyTorch library kernels; `engine.VAEEncoderEngine` is the fp16 sm_100a version
(opt-in `--own_vae_encoder`, 74.8 dB vs this path on seeded weights -- tests/test_03_model_gpu.py).  The CLIP branch of the
front-end runs on this library's own kernels: wiw_b200/clip.py.

Reference: StableVideoDiffusionPipeline._encode_vae_image (DP/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:231-252)
and the SD VAE Encoder (DP/models/autoencoders/vae.py:47-182, blocks in DP/models/unets/unet_2d_blocks.py
DownEncoderBlock2D / UNetMidBlock2D).
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


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
