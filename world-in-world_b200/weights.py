"""Parameter inventory (reference state-dict names and shapes) and seeded random init.

No checkpoints are reachable offline, so benchmarks and parity tests use a seeded
random initialisation laid out under the reference's own state-dict keys
(SURVEY.md section 8a "State-dict contract"); a real `checkpoint-N/unet` safetensors
directory loads through `load_safetensors_dir` into the same dict.

The generator is numpy PCG64 (bit-stable across machines), NOT torch's RNG, so the
build container (golden fixtures) and the GPU box regenerate identical weights.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Tuple

import numpy as np
import torch

from .config import UNetConfig, VAEConfig

Spec = List[Tuple[str, Tuple[int, ...]]]


def _lin(spec: Spec, name: str, cin: int, cout: int, bias: bool = True) -> None:
    spec.append((name + ".weight", (cout, cin)))
    if bias:
        spec.append((name + ".bias", (cout,)))


def _norm(spec: Spec, name: str, c: int) -> None:
    spec.append((name + ".weight", (c,)))
    spec.append((name + ".bias", (c,)))


def _conv(spec: Spec, name: str, cin: int, cout: int, k: Tuple[int, ...]) -> None:
    spec.append((name + ".weight", (cout, cin) + k))
    spec.append((name + ".bias", (cout,)))


def _tsemb(spec: Spec, name: str, cin: int, hid: int, cout: int) -> None:
    _lin(spec, name + ".linear_1", cin, hid)
    _lin(spec, name + ".linear_2", hid, cout)


def _st_resblock(spec: Spec, p: str, cin: int, cout: int, temb: int | None) -> None:
    """SpatioTemporalResBlock (models/resnet.py:635-684)."""
    s = p + ".spatial_res_block"
    _norm(spec, s + ".norm1", cin)
    _conv(spec, s + ".conv1", cin, cout, (3, 3))
    if temb is not None:
        _lin(spec, s + ".time_emb_proj", temb, cout)
    _norm(spec, s + ".norm2", cout)
    _conv(spec, s + ".conv2", cout, cout, (3, 3))
    if cin != cout:
        _conv(spec, s + ".conv_shortcut", cin, cout, (1, 1))
    t = p + ".temporal_res_block"
    _norm(spec, t + ".norm1", cout)
    _conv(spec, t + ".conv1", cout, cout, (3, 1, 1))
    if temb is not None:
        _lin(spec, t + ".time_emb_proj", temb, cout)
    _norm(spec, t + ".norm2", cout)
    _conv(spec, t + ".conv2", cout, cout, (3, 1, 1))
    spec.append((p + ".time_mixer.mix_factor", (1,)))


def _attn(spec: Spec, p: str, c: int, kv: int) -> None:
    _lin(spec, p + ".to_q", c, c, bias=False)
    _lin(spec, p + ".to_k", kv, c, bias=False)
    _lin(spec, p + ".to_v", kv, c, bias=False)
    _lin(spec, p + ".to_out.0", c, c)


def _ff(spec: Spec, p: str, c: int) -> None:
    _lin(spec, p + ".net.0.proj", c, 8 * c)
    _lin(spec, p + ".net.2", 4 * c, c)


def _st_transformer(spec: Spec, p: str, c: int, cross: int) -> None:
    """TransformerSpatioTemporalModel (models/transformers/transformer_temporal.py:219-277)."""
    _norm(spec, p + ".norm", c)
    _lin(spec, p + ".proj_in", c, c)
    b = p + ".transformer_blocks.0"
    _norm(spec, b + ".norm1", c)
    _attn(spec, b + ".attn1", c, c)
    _norm(spec, b + ".norm2", c)
    _attn(spec, b + ".attn2", c, cross)
    _norm(spec, b + ".norm3", c)
    _ff(spec, b + ".ff", c)
    t = p + ".temporal_transformer_blocks.0"
    _norm(spec, t + ".norm_in", c)
    _ff(spec, t + ".ff_in", c)
    _norm(spec, t + ".norm1", c)
    _attn(spec, t + ".attn1", c, c)
    _norm(spec, t + ".norm2", c)
    _attn(spec, t + ".attn2", c, cross)
    _norm(spec, t + ".norm3", c)
    _ff(spec, t + ".ff", c)
    _tsemb(spec, p + ".time_pos_embed", c, 4 * c, c)
    spec.append((p + ".time_mixer.mix_factor", (1,)))
    _lin(spec, p + ".proj_out", c, c)


def unet_param_spec(cfg: UNetConfig) -> Spec:
    """Every parameter of UNetSpatioTemporalConditionModel(action_strategy='micro_cond')."""
    spec: Spec = []
    boc = cfg.block_out_channels
    te = cfg.time_embed_dim
    n = len(boc)
    _conv(spec, "conv_in", cfg.in_channels, boc[0], (3, 3))
    _tsemb(spec, "time_embedding", boc[0], te, te)
    _tsemb(spec, "add_embedding", cfg.projection_class_embeddings_input_dim, te, te)
    out = boc[0]
    for i in range(n):
        cin, out = out, boc[i]
        for j in range(cfg.layers_per_block):
            _st_resblock(spec, f"down_blocks.{i}.resnets.{j}", cin if j == 0 else out, out, te)
            if i < n - 1:
                _st_transformer(spec, f"down_blocks.{i}.attentions.{j}", out, cfg.cross_attention_dim)
        if i < n - 1:
            _conv(spec, f"down_blocks.{i}.downsamplers.0.conv", out, out, (3, 3))
    _st_resblock(spec, "mid_block.resnets.0", boc[-1], boc[-1], te)
    _st_transformer(spec, "mid_block.attentions.0", boc[-1], cfg.cross_attention_dim)
    _st_resblock(spec, "mid_block.resnets.1", boc[-1], boc[-1], te)
    rev = list(reversed(boc))
    out = rev[0]
    for i in range(n):
        prev, out = out, rev[i]
        cin = rev[min(i + 1, n - 1)]
        nl = cfg.layers_per_block + 1
        for j in range(nl):
            skip = cin if j == nl - 1 else out
            rin = prev if j == 0 else out
            _st_resblock(spec, f"up_blocks.{i}.resnets.{j}", rin + skip, out, te)
            if i > 0:
                _st_transformer(spec, f"up_blocks.{i}.attentions.{j}", out, cfg.cross_attention_dim)
        if i < n - 1:
            _conv(spec, f"up_blocks.{i}.upsamplers.0.conv", out, out, (3, 3))
    _norm(spec, "conv_norm_out", boc[0])
    _conv(spec, "conv_out", boc[0], cfg.out_channels, (3, 3))
    _lin(spec, "add_action_proj.proj", cfg.action_input_channel * 12, 256)
    _tsemb(spec, "add_embedding_action", 256, te, te)
    _tsemb(spec, "add_embedding_noise", 256, te, te)
    return spec


def vae_decoder_param_spec(cfg: VAEConfig) -> Spec:
    """Every parameter of TemporalDecoder, keys relative to `decoder.`."""
    spec: Spec = []
    boc = cfg.block_out_channels
    top = boc[-1]
    _conv(spec, "conv_in", cfg.latent_channels, top, (3, 3))
    _st_resblock(spec, "mid_block.resnets.0", top, top, None)
    a = "mid_block.attentions.0"
    _norm(spec, a + ".group_norm", top)
    for nm in ("to_q", "to_k", "to_v", "to_out.0"):
        _lin(spec, f"{a}.{nm}", top, top)
    _st_resblock(spec, "mid_block.resnets.1", top, top, None)
    rev = list(reversed(boc))
    out = rev[0]
    for i in range(len(boc)):
        prev, out = out, rev[i]
        for j in range(cfg.layers_per_block + 1):
            _st_resblock(spec, f"up_blocks.{i}.resnets.{j}", prev if j == 0 else out, out, None)
        if i < len(boc) - 1:
            _conv(spec, f"up_blocks.{i}.upsamplers.0.conv", out, out, (3, 3))
    _norm(spec, "conv_norm_out", boc[0])
    _conv(spec, "conv_out", boc[0], cfg.out_channels, (3, 3))
    _conv(spec, "time_conv_out", cfg.out_channels, cfg.out_channels, (3, 1, 1))
    return spec


def random_init(spec: Spec, seed: int, weight_gain: float = 0.7) -> Dict[str, torch.Tensor]:
    """Seeded fp32 init: matrices/filters N(0, gain^2/fan_in); biases N(0, 0.05^2); norm
    scales 1 + 0.1 N(0,1); mix_factor N(0,1).  One PCG64 stream per tensor (keyed by the
    tensor's index) so the result is independent of generation order."""
    out: Dict[str, torch.Tensor] = {}
    for idx, (name, shape) in enumerate(spec):
        rng = np.random.Generator(np.random.PCG64([seed, idx]))
        x = rng.standard_normal(shape, dtype=np.float32)
        leaf = name.rsplit(".", 1)[-1]
        parent = name.rsplit(".", 2)[-2] if name.count(".") else ""
        is_norm = parent.startswith("norm") or parent in ("group_norm", "conv_norm_out")
        if leaf == "mix_factor":
            pass
        elif is_norm:
            x = (1.0 + 0.1 * x) if leaf == "weight" else 0.1 * x
        elif leaf == "bias":
            x = 0.05 * x
        else:
            fan_in = int(np.prod(shape[1:]))
            x = x * np.float32(weight_gain / np.sqrt(fan_in))
        out[name] = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    return out


def load_safetensors_dir(path: str) -> Dict[str, torch.Tensor]:
    """Read a diffusers-layout `unet/` (or `vae/`) directory -- the same files the reference
    worker loads with from_pretrained (FTsvd/eval_inference.py:116-131)."""
    from safetensors.torch import load_file
    sd: Dict[str, torch.Tensor] = {}
    files = sorted(f for f in os.listdir(path) if f.endswith(".safetensors"))
    if not files:
        raise FileNotFoundError(f"no .safetensors under {path}")
    for f in files:
        sd.update(load_file(os.path.join(path, f)))
    return sd


def check_state_dict(sd: Dict[str, torch.Tensor], spec: Spec, allow_extra: Iterable[str] = ()) -> None:
    """Fail loudly on a missing key or wrong shape (mirrors load_state_dict(strict=True))."""
    names = {n for n, _ in spec}
    missing = [n for n in names if n not in sd]
    bad = [n for n, s in spec if n in sd and tuple(sd[n].shape) != tuple(s)]
    extra = [k for k in sd if k not in names and not any(k.startswith(a) for a in allow_extra)]
    if missing or bad or extra:
        raise KeyError(f"state dict mismatch: missing={missing[:5]} wrong_shape={bad[:5]} unexpected={extra[:5]}")


def random_init_on_device(spec: Spec, seed: int, device, weight_gain: float = 0.7) -> Dict[str, torch.Tensor]:
    """Same distribution as `random_init` but drawn with torch's device generator (seconds instead of
    minutes for the 1.5 B-parameter UNet).  Used by bench.py, where only the statistics matter; parity
    tests use `random_init` so the CPU oracle and the build-container goldens see identical weights."""
    gen = torch.Generator(device=device).manual_seed(seed)
    out: Dict[str, torch.Tensor] = {}
    for name, shape in spec:
        x = torch.randn(shape, generator=gen, device=device, dtype=torch.float32)
        leaf = name.rsplit(".", 1)[-1]
        parent = name.rsplit(".", 2)[-2] if name.count(".") else ""
        is_norm = parent.startswith("norm") or parent in ("group_norm", "conv_norm_out")
        if leaf == "mix_factor":
            pass
        elif is_norm:
            x = (1.0 + 0.1 * x) if leaf == "weight" else 0.1 * x
        elif leaf == "bias":
            x = 0.05 * x
        else:
            x = x * (weight_gain / float(np.sqrt(int(np.prod(shape[1:])))))
        out[name] = x.to(torch.float16)
    return out
