"""CLIP ViT image encoder of the conditioning front-end on the sm_100a kernels.

Reference: StableVideoDiffusionPipeline._encode_image (DP/pipelines/stable_video_diffusion/pipeline_stable_video_diffusion.py:183-229):
the observation AS SENT -> [0,1] -> *2-1 -> _resize_with_antialiasing to 224x224 (:643-747) -> (x+1)/2 -> CLIP mean/std
normalisation (feature_extractor, do_resize / do_center_crop / do_rescale all off) -> image_encoder(...).image_embeds.
The encoder is third-party code (transformers CLIPVisionModelWithProjection); its checkpoint layout (state-dict names) is the
boundary here, and parity is checked against the installed transformers module on the same weights (tests/test_06_clip_gpu.py).

Design, not a port:
  * the antialiased resize is LINEAR, so blur (reflect padding) and bicubic (align_corners=True, A = -0.75, clamped taps) are
    composed on the host into ONE banded operator per axis; because both kernels sum to one, the pipeline's `*2-1` / `(x+1)/2`
    round trip cancels exactly and the CLIP normalisation is applied to the operator's output: one pair of launches takes the
    uint8 observation to the normalised fp16 224x224 planes (wiw_resize_separable_f32);
  * patch embedding = im2col (wiw_patchify) + the implicit-GEMM kernel, with the position embedding added in its epilogue as a
    per-row table; q/k/v are one fused GEMM; residual adds ride in the GEMM epilogues; attention over 257 tokens x 16 heads of 80
    is a small shared-memory kernel (wiw_attention_small); the MLP activation (GELU or quick-GELU) is a GEMM epilogue mode.
"""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import Dict, List, Tuple

import numpy as np
import torch

from . import lib as L
from . import ops

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


# ----------------------------------------------------------------------------- antialiased resize as one operator per axis
def _cubic_taps(t: np.ndarray, a: float = -0.75) -> np.ndarray:
    """Cubic-convolution weights of the four taps i0-1 .. i0+2 at fractional offset t (ATen upsample_bicubic2d)."""
    def c1(x):
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0

    def c2(x):
        return ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return np.stack([c2(t + 1.0), c1(t), c1(1.0 - t), c2(2.0 - t)], axis=-1)


@lru_cache(maxsize=32)
def antialias_operator(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray]:
    """(bounds int32 [out, 2] = (first source index, taps), weights fp32 [out, ksize]) of
    bicubic(align_corners=True)( gaussian_blur_reflect(x) ) along one axis, as the pipeline parametrises it (:643-667):
    sigma = max((in/out - 1) / 2, 0.001), kernel size = max(int(4 sigma), 3) made odd."""
    factor = in_size / out_size
    sigma = max((factor - 1.0) / 2.0, 0.001)
    ks = int(max(2.0 * 2.0 * sigma, 3))
    ks += (ks % 2 == 0)
    x = np.arange(ks, dtype=np.float64) - ks // 2
    g = np.exp(-(x ** 2) / (2.0 * sigma ** 2))
    g /= g.sum()
    p = ks // 2
    blur = np.zeros((in_size, in_size), np.float64)                      # reflect padding folded into the matrix
    for i in range(in_size):
        for t in range(ks):
            j = i + t - p
            if j < 0:
                j = -j
            elif j >= in_size:
                j = 2 * (in_size - 1) - j
            blur[i, j] += g[t]
    scale = np.float32(in_size - 1) / np.float32(out_size - 1) if out_size > 1 else np.float32(0)
    src = (np.arange(out_size, dtype=np.float32) * scale).astype(np.float32)   # ATen computes the source index in fp32
    i0 = np.floor(src).astype(np.int64)
    taps = _cubic_taps((src - i0).astype(np.float64))
    op = np.zeros((out_size, in_size), np.float64)
    for o in range(out_size):
        for k in range(4):
            j = min(max(int(i0[o]) - 1 + k, 0), in_size - 1)              # border taps are clamped, not dropped
            op[o] += taps[o, k] * blur[j]
    nz = np.abs(op) > 0
    first = nz.argmax(axis=1)
    last = in_size - 1 - nz[:, ::-1].argmax(axis=1)
    ksize = int((last - first + 1).max())
    bounds = np.stack([first, last - first + 1], axis=1).astype(np.int32)
    w = np.zeros((out_size, ksize), np.float32)
    for o in range(out_size):
        w[o, : bounds[o, 1]] = op[o, first[o]: last[o] + 1]
    return bounds, w


class ClipPreprocessor:
    """uint8 observations [B, 3, H, W] on the device -> CLIP-normalised fp16 [B, 3, size, size]."""

    def __init__(self, device, size: int = 224):
        self.device, self.size = torch.device(device), size
        self.mean = torch.tensor(CLIP_MEAN, dtype=torch.float32, device=self.device)
        self.inv_std = (1.0 / torch.tensor(CLIP_STD, dtype=torch.float64)).to(torch.float32).to(self.device)
        self._ops: Dict[Tuple[int, int], tuple] = {}

    def _operator(self, n_in: int):
        key = (n_in, self.size)
        if key not in self._ops:
            b, w = antialias_operator(n_in, self.size)
            self._ops[key] = (torch.from_numpy(b.copy()).to(self.device), torch.from_numpy(w.copy()).to(self.device), w.shape[1])
        return self._ops[key]

    def __call__(self, images_u8: torch.Tensor) -> torch.Tensor:
        lib = L.require_device()
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[1] != 3 or not images_u8.is_contiguous():
            raise L.WiwError(f"expected contiguous uint8 [B, 3, H, W], got {images_u8.dtype} {tuple(images_u8.shape)}")
        B, C, H, W = images_u8.shape
        bh, wh, kh = self._operator(W)
        bv, wv, kv = self._operator(H)
        tmp = torch.empty(B * C, H, self.size, dtype=torch.float32, device=images_u8.device)
        out = torch.empty(B, C, self.size, self.size, dtype=torch.float16, device=images_u8.device)
        L.check(lib.wiw_resize_separable_f32(images_u8.data_ptr(), tmp.data_ptr(), out.data_ptr(), B * C, C, H, W, self.size, self.size,
                                             bh.data_ptr(), wh.data_ptr(), kh, bv.data_ptr(), wv.data_ptr(), kv,
                                             self.mean.data_ptr(), self.inv_std.data_ptr(), L.stream_ptr()))
        return out


# ----------------------------------------------------------------------------- the ViT
@dataclass
class _Layer:
    ln1: ops.NormWeight
    qkv: ops.GemmWeight
    out: ops.GemmWeight
    ln2: ops.NormWeight
    fc1: ops.GemmWeight
    fc2: ops.GemmWeight


ACT = {"gelu": 2, "quick_gelu": 3}


class ClipVisionEngine:
    """transformers.CLIPVisionModelWithProjection on the sm_100a kernels; `sd` is its state dict, `cfg` its config (hidden_size,
    num_attention_heads, num_hidden_layers, patch_size, image_size, hidden_act, layer_norm_eps).  hidden_size and
    intermediate_size must be multiples of 64 (ViT-H/14: 1280 / 5120)."""

    def __init__(self, sd: Dict[str, torch.Tensor], cfg, device="cuda"):
        self.device = torch.device(device)
        self.hidden, self.heads = int(cfg.hidden_size), int(cfg.num_attention_heads)
        self.d = self.hidden // self.heads
        self.ps, self.size = int(cfg.patch_size), int(cfg.image_size)
        self.grid = self.size // self.ps
        self.eps = float(cfg.layer_norm_eps)
        if cfg.hidden_act not in ACT:
            raise ValueError(f"unsupported CLIP activation {cfg.hidden_act!r}")
        self.act = ACT[cfg.hidden_act]
        if self.hidden % 64 or int(cfg.intermediate_size) % 64 or self.d % 2 or self.d > 128:
            raise ValueError("ClipVisionEngine needs hidden / intermediate sizes that are multiples of 64 and an even head_dim <= 128")
        dev, v = self.device, "vision_model."
        f = lambda k: sd[k].float()
        self.pre = ClipPreprocessor(dev, self.size)
        self.patch = ops.prep_linear(f(v + "embeddings.patch_embedding.weight").flatten(1), None, dev)
        pos = f(v + "embeddings.position_embedding.weight")
        self.pos_patches = pos[1:].to(torch.float16).to(dev).contiguous()                       # epilogue table of the patch GEMM
        self.cls_row = (f(v + "embeddings.class_embedding").to(torch.float16) + pos[0].to(torch.float16)).to(dev)   # fp16 add, as the fp16 module does
        self.pre_ln = ops.prep_norm(f(v + "pre_layrnorm.weight"), f(v + "pre_layrnorm.bias"), dev)
        self.layers: List[_Layer] = []
        for i in range(int(cfg.num_hidden_layers)):
            p = f"{v}encoder.layers.{i}."
            wqkv = torch.cat([f(p + "self_attn.q_proj.weight"), f(p + "self_attn.k_proj.weight"), f(p + "self_attn.v_proj.weight")])
            bqkv = torch.cat([f(p + "self_attn.q_proj.bias"), f(p + "self_attn.k_proj.bias"), f(p + "self_attn.v_proj.bias")])
            self.layers.append(_Layer(
                ops.prep_norm(f(p + "layer_norm1.weight"), f(p + "layer_norm1.bias"), dev), ops.prep_linear(wqkv, bqkv, dev),
                ops.prep_linear(f(p + "self_attn.out_proj.weight"), f(p + "self_attn.out_proj.bias"), dev),
                ops.prep_norm(f(p + "layer_norm2.weight"), f(p + "layer_norm2.bias"), dev),
                ops.prep_linear(f(p + "mlp.fc1.weight"), f(p + "mlp.fc1.bias"), dev),
                ops.prep_linear(f(p + "mlp.fc2.weight"), f(p + "mlp.fc2.bias"), dev)))
        self.post_ln = ops.prep_norm(f(v + "post_layernorm.weight"), f(v + "post_layernorm.bias"), dev)
        self.proj = ops.prep_linear(f("visual_projection.weight"), None, dev)
        self.proj_dim = int(sd["visual_projection.weight"].shape[0])

    @classmethod
    def from_module(cls, module, device="cuda") -> "ClipVisionEngine":
        return cls({k: t.detach() for k, t in module.state_dict().items()}, module.config, device)

    @torch.no_grad()
    def encode_pixels(self, pixels: torch.Tensor) -> torch.Tensor:
        """CLIP-normalised fp16 pixels [B, 3, size, size] -> image_embeds fp16 [B, projection_dim]."""
        lib = L.require_device()
        B = pixels.shape[0]
        G, T, H = self.grid, self.grid * self.grid + 1, self.hidden
        cols = torch.empty(B * G * G, self.patch.k, dtype=torch.float16, device=self.device)
        L.check(lib.wiw_patchify(pixels.data_ptr(), cols.data_ptr(), B, 3, G, self.ps, self.patch.k, L.stream_ptr()))
        hs = torch.empty(B, T, H, dtype=torch.float16, device=self.device)
        hs[:, 0] = self.cls_row
        for b in range(B):                                       # patch rows of image b land behind its class token
            ops.linear(cols[b * G * G:(b + 1) * G * G], self.patch, tab=self.pos_patches, tab_div=1, tab_mod=G * G, out=hs[b, 1:])
        x = ops.layer_norm(hs.view(B * T, H), self.pre_ln, self.eps)
        for ly in self.layers:
            qkv = ops.linear(ops.layer_norm(x, ly.ln1, self.eps), ly.qkv)
            att = torch.empty(B * T, H, dtype=torch.float16, device=self.device)
            L.check(lib.wiw_attention_small(qkv.data_ptr(), att.data_ptr(), B, T, self.heads, self.d, float(self.d) ** -0.5, L.stream_ptr()))
            x = ops.linear(att, ly.out, res=x)
            u = ops.linear(ops.layer_norm(x, ly.ln2, self.eps), ly.fc1, act=self.act)
            x = ops.linear(u, ly.fc2, res=x)
        pooled = x.view(B, T, H)[:, 0].contiguous()
        emb = ops.linear(ops.layer_norm(pooled, self.post_ln, self.eps), self.proj)
        return emb[:, : self.proj_dim]

    @torch.no_grad()
    def __call__(self, images_u8: torch.Tensor) -> torch.Tensor:
        """uint8 observations [B, 3, H, W] (any size) on the device -> image embedding fp16 [B, 1, projection_dim] (pipeline :183-216)."""
        return self.encode_pixels(self.pre(images_u8)).unsqueeze(1)
