"""Plain-PyTorch CPU restatement of the reference's SVD-dagger hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Functional code over a flat
state dict that uses the reference's own parameter names, NCHW tensors, fp32 by
default.  Every function cites the reference file:line it restates
(paths relative to /root/reference; DP = FTsvd/diffusers-private/diffusers).

Pinned against outputs of the vendored fork itself: tests/golden/*.npz, made by
oracle/make_golden.py, checked by tests/test_oracle_golden.py.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------- configs
@dataclass
class UNetCfg:
    """Constructor arguments of UNetSpatioTemporalConditionModel
    (DP/models/unets/unet_spatio_temporal_condition.py:72-97) that the path uses."""
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    cross_attention_dim: int = 1024
    addition_time_embed_dim: int = 256
    num_frames: int = 14
    action_input_channel: int = 14


@dataclass
class VAECfg:
    """TemporalDecoder config (DP/models/autoencoders/autoencoder_kl_temporal_decoder.py:29-36)."""
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    scaling_factor: float = 0.18215


def _sub(sd: SD, prefix: str) -> SD:
    p = prefix + "."
    return {k[len(p):]: v for k, v in sd.items() if k.startswith(p)}


# --------------------------------------------------------------------------- small ops
def linear(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def group_norm(sd: SD, name: str, x: torch.Tensor, eps: float, groups: int = 32) -> torch.Tensor:
    return F.group_norm(x, groups, sd[name + ".weight"], sd[name + ".bias"], eps)


def layer_norm(sd: SD, name: str, x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def get_timestep_embedding(timesteps: torch.Tensor, dim: int) -> torch.Tensor:
    """DP/models/embeddings.py:27-78 with flip_sin_to_cos=True, downscale_freq_shift=0,
    scale=1 (how every Timesteps module on the path is built: unet :147,151;
    transformer_temporal.py:268)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32)
    exponent = exponent / (half - 0.0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)       # flip -> [cos, sin]
    return emb


def timestep_embedding_mlp(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """TimestepEmbedding.forward, DP/models/embeddings.py:800-816 (act = SiLU)."""
    return linear(sd, name + ".linear_2", F.silu(linear(sd, name + ".linear_1", x)))


def action_embedder(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """ActionEmbedder_.forward, DP/models/embeddings.py:912-945: 12 Fourier features per
    element, order cos,sin of k*x for k in 1,2,4,6,8,10, flattened (elem-major), Linear."""
    bs, n, m = x.shape
    feats = torch.stack([
        torch.cos(x), torch.sin(x), torch.cos(2.0 * x), torch.sin(2.0 * x),
        torch.cos(4.0 * x), torch.sin(4.0 * x), torch.cos(x * 6.0), torch.sin(x * 6.0),
        torch.cos(x * 8.0), torch.sin(x * 8.0), torch.cos(x * 10.0), torch.sin(x * 10.0),
    ], dim=-1).view(bs, n, m * 12)
    return linear(sd, name + ".proj", feats.to(sd[name + ".proj.weight"].dtype))


def action_ids_idx_encode(actions: torch.Tensor) -> torch.Tensor:
    """utils/svd_utils.py:594-632: cumulative matrix A[b,i,j] = act[b,j] for j<=i else 0,
    with column 0 forced to 'stop' (=4).  Integer semantics; returns int64 [b,L,L]."""
    assert actions.dim() == 2
    b, L = actions.shape
    out = torch.zeros(b, L, L, dtype=torch.long)
    for bi in range(b):
        frame = torch.zeros(L, dtype=torch.long)
        for i in range(L):
            frame[i] = 4 if i == 0 else int(actions[bi, i])
            out[bi, i] = frame
    return out


# --------------------------------------------------------------------------- attention
def attention(sd: SD, name: str, x: torch.Tensor, ctx: Optional[torch.Tensor], heads: int) -> torch.Tensor:
    """Attention + AttnProcessor2_0 (DP/models/attention_processor.py:2321-2403) for the
    transformer blocks: no q/k/v bias, to_out bias, scale 1/sqrt(head_dim), no mask."""
    kv = x if ctx is None else ctx
    q = linear(sd, name + ".to_q", x)
    k = linear(sd, name + ".to_k", kv)
    v = linear(sd, name + ".to_v", kv)
    B, S, C = q.shape
    d = C // heads
    q = q.view(B, -1, heads, d).transpose(1, 2)
    k = k.view(B, -1, heads, d).transpose(1, 2)
    v = v.view(B, -1, heads, d).transpose(1, 2)
    o = F.scaled_dot_product_attention(q, k, v)
    o = o.transpose(1, 2).reshape(B, -1, C)
    return linear(sd, name + ".to_out.0", o)


def feed_forward(sd: SD, name: str, x: torch.Tensor) -> torch.Tensor:
    """FeedForward with GEGLU (DP/models/attention.py:1185-1243, activations.py:93-123):
    proj -> chunk(value, gate) -> value * gelu_erf(gate) -> Linear."""
    h = linear(sd, name + ".net.0.proj", x)
    val, gate = h.chunk(2, dim=-1)
    return linear(sd, name + ".net.2", val * F.gelu(gate))


def basic_transformer_block(sd: SD, x: torch.Tensor, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """BasicTransformerBlock.forward, layer_norm path (DP/models/attention.py:474-582)."""
    x = attention(sd, "attn1", layer_norm(sd, "norm1", x), None, heads) + x
    x = attention(sd, "attn2", layer_norm(sd, "norm2", x), ctx, heads) + x
    x = feed_forward(sd, "ff", layer_norm(sd, "norm3", x)) + x
    return x


def temporal_transformer_block(sd: SD, x: torch.Tensor, num_frames: int, ctx: torch.Tensor, heads: int) -> torch.Tensor:
    """TemporalBasicTransformerBlock.forward (DP/models/attention.py:702-762); is_res=True."""
    bf, s, c = x.shape
    b = bf // num_frames
    x = x[None, :].reshape(b, num_frames, s, c).permute(0, 2, 1, 3).reshape(b * s, num_frames, c)
    residual = x
    x = feed_forward(sd, "ff_in", layer_norm(sd, "norm_in", x)) + residual
    x = attention(sd, "attn1", layer_norm(sd, "norm1", x), None, heads) + x
    x = attention(sd, "attn2", layer_norm(sd, "norm2", x), ctx, heads) + x
    x = feed_forward(sd, "ff", layer_norm(sd, "norm3", x)) + x
    x = x[None, :].reshape(b, s, num_frames, c).permute(0, 2, 1, 3).reshape(b * num_frames, s, c)
    return x


def transformer_spatio_temporal(sd: SD, x: torch.Tensor, enc: torch.Tensor, num_frames: int, heads: int) -> torch.Tensor:
    """TransformerSpatioTemporalModel.forward (DP/models/transformers/transformer_temporal.py:279-382),
    one layer; image_only_indicator == 0 so alpha = sigmoid(mix_factor) (resnet.py:756-797)."""
    bf, c, h, w = x.shape
    b = bf // num_frames
    tc = enc[None, :].reshape(b, num_frames, -1, enc.shape[-1])[:, 0]
    tc = tc[:, None].broadcast_to(b, h * w, tc.shape[-2], tc.shape[-1]).reshape(b * h * w, -1, enc.shape[-1])
    residual = x
    x = group_norm(sd, "norm", x, 1e-6)
    x = x.permute(0, 2, 3, 1).reshape(bf, h * w, c)
    x = linear(sd, "proj_in", x)
    fidx = torch.arange(num_frames).repeat(b, 1).reshape(-1)
    t_emb = get_timestep_embedding(fidx, c).to(x.dtype)
    emb = timestep_embedding_mlp(sd, "time_pos_embed", t_emb)[:, None, :]
    x = basic_transformer_block(_sub(sd, "transformer_blocks.0"), x, enc, heads)
    x_mix = x + emb
    x_mix = temporal_transformer_block(_sub(sd, "temporal_transformer_blocks.0"), x_mix, num_frames, tc, heads)
    alpha = torch.sigmoid(sd["time_mixer.mix_factor"]).to(x.dtype)
    x = alpha * x + (1.0 - alpha) * x_mix
    x = linear(sd, "proj_out", x)
    x = x.reshape(bf, h, w, c).permute(0, 3, 1, 2).contiguous()
    return x + residual


# --------------------------------------------------------------------------- resnets
def resnet_block_2d(sd: SD, x: torch.Tensor, temb: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """ResnetBlock2D.forward, 'default' time-embedding norm (DP/models/resnet.py:320-373)."""
    h = F.silu(group_norm(sd, "norm1", x, eps))
    h = F.conv2d(h, sd["conv1.weight"], sd["conv1.bias"], padding=1)
    if temb is not None:
        h = h + linear(sd, "time_emb_proj", F.silu(temb))[:, :, None, None]
    h = F.silu(group_norm(sd, "norm2", h, eps))
    h = F.conv2d(h, sd["conv2.weight"], sd["conv2.bias"], padding=1)
    if "conv_shortcut.weight" in sd:
        x = F.conv2d(x, sd["conv_shortcut.weight"], sd["conv_shortcut.bias"])
    return x + h


def temporal_resnet_block(sd: SD, x: torch.Tensor, temb: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """TemporalResnetBlock.forward on [B,C,F,h,w] (DP/models/resnet.py:604-631)."""
    h = F.silu(group_norm(sd, "norm1", x, eps))
    h = F.conv3d(h, sd["conv1.weight"], sd["conv1.bias"], padding=(1, 0, 0))
    if temb is not None:
        t = linear(sd, "time_emb_proj", F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
        h = h + t
    h = F.silu(group_norm(sd, "norm2", h, eps))
    h = F.conv3d(h, sd["conv2.weight"], sd["conv2.bias"], padding=(1, 0, 0))
    return x + h


def spatio_temporal_res_block(sd: SD, x: torch.Tensor, temb: Optional[torch.Tensor], num_frames: int,
                              eps: float, temporal_eps: Optional[float] = None, switch: bool = False) -> torch.Tensor:
    """SpatioTemporalResBlock.forward (DP/models/resnet.py:686-716) + AlphaBlender (:784-797).
    `switch` = switch_spatial_to_temporal_mix (temporal VAE decoder)."""
    x = resnet_block_2d(_sub(sd, "spatial_res_block"), x, temb, eps)
    bf, c, h, w = x.shape
    b = bf // num_frames
    xs = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    t = temb.reshape(b, num_frames, -1) if temb is not None else None
    xt = temporal_resnet_block(_sub(sd, "temporal_res_block"), xs, t, temporal_eps if temporal_eps is not None else eps)
    alpha = torch.sigmoid(sd["time_mixer.mix_factor"]).to(x.dtype)
    if switch:
        alpha = 1.0 - alpha
    out = alpha * xs + (1.0 - alpha) * xt
    return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# --------------------------------------------------------------------------- UNet
def unet_embeddings(sd: SD, cfg: UNetCfg, timestep: torch.Tensor, added_time_ids: torch.Tensor,
                    added_action_ids: torch.Tensor, nb: int, dtype: torch.dtype) -> torch.Tensor:
    """micro_cond embedding, unet :432-487 under no_grad (inference): returns emb [nb*F, 1280].
    Keeps the reference's `repeat(1,2,1)` action-row behaviour (valid pairing only for B==1)."""
    F_ = cfg.num_frames
    t = timestep.reshape(-1)[:1].expand(nb)
    t_emb = get_timestep_embedding(t, cfg.block_out_channels[0]).to(dtype)
    emb_only_time = timestep_embedding_mlp(sd, "time_embedding", t_emb).repeat_interleave(F_, dim=0)
    act = timestep_embedding_mlp(sd, "add_embedding_action", action_embedder(sd, "add_action_proj", added_action_ids))
    act = act.repeat(1, 2, 1)                                   # unet :479-480 (no_grad branch)
    act = act.reshape(-1, act.shape[-1])
    emb = emb_only_time + act.to(dtype)
    noise = get_timestep_embedding(added_time_ids[:, -1].flatten(), cfg.addition_time_embed_dim).to(dtype)
    noise = timestep_embedding_mlp(sd, "add_embedding_noise", noise).repeat_interleave(F_, dim=0)
    return emb + noise


def unet_forward(sd: SD, cfg: UNetCfg, sample: torch.Tensor, timestep, enc: torch.Tensor,
                 added_time_ids: torch.Tensor, added_action_ids: torch.Tensor,
                 taps: Optional[dict] = None) -> torch.Tensor:
    """UNetSpatioTemporalConditionModel.forward (unet :402-575), micro_cond / navigation.
    sample [nb,F,8,h,w]; enc [nb,1,cross]; added_time_ids [nb,3]; added_action_ids [B,F,A].
    `taps` (optional dict) collects named intermediate activations for per-block tests."""
    nb, F_ = sample.shape[:2]
    dtype = sample.dtype
    timestep = torch.as_tensor(timestep, dtype=torch.float32)
    emb = unet_embeddings(sd, cfg, timestep, added_time_ids, added_action_ids, nb, dtype)
    x = sample.flatten(0, 1)
    enc_rep = enc.repeat_interleave(F_, dim=0)
    x = F.conv2d(x, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    if taps is not None:
        taps["emb"] = emb
        taps["conv_in"] = x
    skips = [x]
    nblk = len(cfg.block_out_channels)
    heads = cfg.num_attention_heads
    # down (unet :514-531; block eps: CrossAttn 1e-6 unet_3d_blocks.py:1308, Down 1e-5 :1202)
    for i in range(nblk):
        has_attn = i < nblk - 1
        eps = 1e-6 if has_attn else 1e-5
        for j in range(cfg.layers_per_block):
            x = spatio_temporal_res_block(_sub(sd, f"down_blocks.{i}.resnets.{j}"), x, emb, F_, eps)
            if has_attn:
                x = transformer_spatio_temporal(_sub(sd, f"down_blocks.{i}.attentions.{j}"), x, enc_rep, F_, heads[i])
            skips.append(x)
            if taps is not None:
                taps[f"down.{i}.{j}"] = x
        if i < nblk - 1:
            x = F.conv2d(x, sd[f"down_blocks.{i}.downsamplers.0.conv.weight"],
                         sd[f"down_blocks.{i}.downsamplers.0.conv.bias"], stride=2, padding=1)
            skips.append(x)
    # mid (unet_3d_blocks.py:1109-1180; eps 1e-5 :1069,1100)
    x = spatio_temporal_res_block(_sub(sd, "mid_block.resnets.0"), x, emb, F_, 1e-5)
    x = transformer_spatio_temporal(_sub(sd, "mid_block.attentions.0"), x, enc_rep, F_, heads[-1])
    x = spatio_temporal_res_block(_sub(sd, "mid_block.resnets.1"), x, emb, F_, 1e-5)
    if taps is not None:
        taps["mid"] = x
    # up (unet :543-562).  get_up_block does not forward resnet_eps, so the class default
    # 1e-6 applies to both UpBlockSpatioTemporal (:1441) and CrossAttnUp (:1531).
    rheads = list(reversed(heads))
    for i in range(nblk):
        has_attn = i > 0
        for j in range(cfg.layers_per_block + 1):
            x = torch.cat([x, skips.pop()], dim=1)
            x = spatio_temporal_res_block(_sub(sd, f"up_blocks.{i}.resnets.{j}"), x, emb, F_, 1e-6)
            if has_attn:
                x = transformer_spatio_temporal(_sub(sd, f"up_blocks.{i}.attentions.{j}"), x, enc_rep, F_, rheads[i])
            if taps is not None:
                taps[f"up.{i}.{j}"] = x
        if i < nblk - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")       # upsampling.py:169
            x = F.conv2d(x, sd[f"up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(group_norm(sd, "conv_norm_out", x, 1e-5))
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    return x.reshape(nb, F_, *x.shape[1:])


# --------------------------------------------------------------------------- temporal VAE decoder
def vae_attention(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """Attention with group_norm, bias=True, residual_connection, heads = C // head_dim = 1
    (unet_3d_blocks.py:958-969; attention_processor.py:2321-2403 4-D input path)."""
    b, c, h, w = x.shape
    residual = x
    t = x.view(b, c, h * w).transpose(1, 2)
    t = F.group_norm(t.transpose(1, 2), 32, sd["group_norm.weight"], sd["group_norm.bias"], 1e-6).transpose(1, 2)
    q, k, v = linear(sd, "to_q", t), linear(sd, "to_k", t), linear(sd, "to_v", t)
    o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
    o = linear(sd, "to_out.0", o)
    return o.transpose(-1, -2).reshape(b, c, h, w) + residual


def vae_decode(sd: SD, cfg: VAECfg, z: torch.Tensor, num_frames: int) -> torch.Tensor:
    """TemporalDecoder.forward (autoencoder_kl_temporal_decoder.py:87-161); sd keys are
    relative to `decoder.`; z is already divided by scaling_factor (pipeline :284)."""
    boc = cfg.block_out_channels
    kw = dict(temb=None, num_frames=num_frames, eps=1e-6, temporal_eps=1e-5, switch=True)
    x = F.conv2d(z, sd["conv_in.weight"], sd["conv_in.bias"], padding=1)
    x = spatio_temporal_res_block(_sub(sd, "mid_block.resnets.0"), x, **kw)
    x = vae_attention(_sub(sd, "mid_block.attentions.0"), x)
    x = spatio_temporal_res_block(_sub(sd, "mid_block.resnets.1"), x, **kw)
    for i in range(len(boc)):
        for j in range(cfg.layers_per_block + 1):
            x = spatio_temporal_res_block(_sub(sd, f"up_blocks.{i}.resnets.{j}"), x, **kw)
        if i < len(boc) - 1:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
            x = F.conv2d(x, sd[f"up_blocks.{i}.upsamplers.0.conv.weight"],
                         sd[f"up_blocks.{i}.upsamplers.0.conv.bias"], padding=1)
    x = F.silu(group_norm(sd, "conv_norm_out", x, 1e-6))
    x = F.conv2d(x, sd["conv_out.weight"], sd["conv_out.bias"], padding=1)
    bf, c, h, w = x.shape
    b = bf // num_frames
    x = x[None, :].reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
    x = F.conv3d(x, sd["time_conv_out.weight"], sd["time_conv_out.bias"], padding=(1, 0, 0))
    return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# --------------------------------------------------------------------------- scheduler
class EulerSchedule:
    """EulerDiscreteScheduler with the SVD config (scheduling_euler_discrete.py):
    scaled_linear betas 0.00085..0.012, v_prediction, continuous timesteps, Karras sigmas
    sigma_min 0.002 / sigma_max 700, 'leading' spacing, steps_offset 1.

    set_timesteps :319-450, _convert_to_karras :476-499: Karras ramp in float64 numpy,
    append 0, cast fp32; timesteps = 0.25*log(sigma) per element in torch fp32 (:444)."""

    def __init__(self, num_inference_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0):
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho = sigma_min ** (1 / rho)
        max_inv_rho = sigma_max ** (1 / rho)
        sig = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
        sig = np.concatenate([sig, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig).to(torch.float32)
        self.timesteps = torch.Tensor([0.25 * s.log() for s in self.sigmas[:-1]])
        self.num_inference_steps = num_inference_steps

    @property
    def init_noise_sigma(self) -> torch.Tensor:
        """:262-268: 'leading' spacing -> sqrt(sigma_max^2 + 1)."""
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def index_for_timestep(self, t: torch.Tensor) -> int:
        """:554-566 -- position of t in timesteps (second match if duplicated)."""
        idx = (self.timesteps == t).nonzero()
        pos = 1 if len(idx) > 1 else 0
        return int(idx[pos].item())

    def scale_model_input(self, sample: torch.Tensor, i: int) -> torch.Tensor:
        """:295-317."""
        sigma = self.sigmas[i]
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output: torch.Tensor, i: int, sample: torch.Tensor) -> torch.Tensor:
        """:576-684, gamma = 0, v_prediction; fp32 maths, result cast to model_output.dtype."""
        sample = sample.to(torch.float32)
        sigma = self.sigmas[i]
        pred_original = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        derivative = (sample - pred_original) / sigma
        dt = self.sigmas[i + 1] - sigma
        prev = sample + derivative * dt
        return prev.to(model_output.dtype)


# --------------------------------------------------------------------------- latents / rollout
def sample_latent_noise(noise_all: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """pipeline :750-786 with the randn draw factored out: noise_all [B,F,4,h,w] is the
    raw draw; frames f>=1 with action 2 / 3 become roll(prev, +w/16 / -w/16) along width
    (rotate_by_degrees :795-823; requires w % 16 == 0)."""
    out = noise_all.clone()
    B, F_ = actions.shape
    w = out.shape[-1]
    assert w % 16 == 0, "Width must be divisible by 16 for 22.5-degree rotation."
    shift = w // 16
    for b in range(B):
        for i in range(1, F_):
            a = int(actions[b, i])
            if a == 2:
                out[b, i] = torch.roll(out[b, i - 1].clone(), shifts=shift, dims=-1)
            elif a == 3:
                out[b, i] = torch.roll(out[b, i - 1].clone(), shifts=-shift, dims=-1)
    return out


def denoise_rollout(unet_sd: SD, cfg: UNetCfg, latents: torch.Tensor, image_latents: torch.Tensor,
                    image_embeddings: torch.Tensor, added_action_ids: torch.Tensor, num_steps: int,
                    fps: int = 7, motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                    min_guidance: float = 1.0, max_guidance: float = 3.0,
                    trace: Optional[list] = None) -> torch.Tensor:
    """Pipeline steps 5-9 (pipeline :545-622) for B=1 with CFG.
    latents: [1,F,4,h,w] ALREADY multiplied by init_noise_sigma (:359-361);
    image_latents [1,4,h,w] (cond row; CFG zero row is added here, :246-249);
    image_embeddings [1,1,D] (cond row; zero row added, :221-227)."""
    F_ = cfg.num_frames
    dtype = latents.dtype
    sched = EulerSchedule(num_steps)
    il = torch.cat([torch.zeros_like(image_latents), image_latents]).unsqueeze(1).repeat(1, F_, 1, 1, 1)
    ie = torch.cat([torch.zeros_like(image_embeddings), image_embeddings])
    ids = torch.tensor([[fps - 1, motion_bucket_id, noise_aug_strength]], dtype=dtype).repeat(2, 1)
    gs = torch.linspace(min_guidance, max_guidance, F_).unsqueeze(0).to(dtype)[:, :, None, None, None]
    for i, t in enumerate(sched.timesteps):
        x_in = torch.cat([latents] * 2)
        x_in = sched.scale_model_input(x_in, i)
        x_in = torch.cat([x_in, il], dim=2)
        v = unet_forward(unet_sd, cfg, x_in, t, ie, ids, added_action_ids)
        vu, vc = v.chunk(2)
        v = vu + gs * (vc - vu)
        latents = sched.step(v, i, latents)
        if trace is not None:
            trace.append(latents.clone())
    return latents


def decode_and_postprocess(vae_sd: SD, vcfg: VAECfg, latents: torch.Tensor) -> torch.Tensor:
    """decode_latents (pipeline :282-309) + VaeImageProcessor.postprocess 'pt' denormalise
    (DP/image_processor.py:239): returns float32 [B,F,3,H,W] in [0,1]."""
    B, F_ = latents.shape[:2]
    z = latents.flatten(0, 1)
    z = 1 / vcfg.scaling_factor * z
    frames = vae_decode(vae_sd, vcfg, z, F_)
    frames = frames.reshape(B, F_, *frames.shape[1:]).float()
    return (frames / 2 + 0.5).clamp(0, 1)


def frames_to_uint8(frames01: torch.Tensor) -> np.ndarray:
    """numpy_to_pil quantisation (DP/image_processor.py:147): (x*255).round().astype(uint8)."""
    return (frames01.numpy() * 255).round().astype(np.uint8)


def psnr(a: np.ndarray, b: np.ndarray) -> float:
    """evaluation/FVD/calculate_psnr.py:6-15 on [0,1] images."""
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    if mse < 1e-10:
        return 100.0
    return 20 * math.log10(1 / math.sqrt(mse))
