"""Host orchestration of the SVD-dagger UNet and temporal-VAE decoder on the sm_100a kernels.

Python only sequences kernel launches (one torch stream, no syncs, CUDA-graph capturable);
all arithmetic happens in libwiw_b200.so.  Activations are fp16 channels-last:
a "(b f) c h w" tensor of the reference lives here as [b*f, h, w, c] / [b, f, h*w, c].

Reference structure followed (DP = FTsvd/diffusers-private/diffusers):
  UNetSpatioTemporalConditionModel.forward   DP/models/unets/unet_spatio_temporal_condition.py:402-575
  SpatioTemporalResBlock                     DP/models/resnet.py:635-716 (+ ResnetBlock2D :320-373,
                                             TemporalResnetBlock :604-631, AlphaBlender :784-797)
  TransformerSpatioTemporalModel             DP/models/transformers/transformer_temporal.py:279-382
  BasicTransformerBlock / Temporal...        DP/models/attention.py:474-582 / :702-762
  TemporalDecoder                            DP/models/autoencoders/autoencoder_kl_temporal_decoder.py:87-161

Algebraic simplification used (exact): cross-attention over a length-1 context
is softmax over one key == 1, so attn2(x, ctx) = to_out(to_v(ctx)) for every token; it becomes a
per-sample vector added in the epilogue of the preceding projection.  Valid for num_past_obs == 1
(the WM-server setting); a longer context raises.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from .config import UNetConfig, VAEConfig

SD = Dict[str, torch.Tensor]


def sinusoid(values: torch.Tensor, dim: int) -> torch.Tensor:
    """get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0), DP/models/embeddings.py:27-78,
    evaluated with the same torch fp32 ops as the reference (host side; tables are tiny)."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half
    emb = values.reshape(-1)[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


@dataclass
class _Res:
    norm1: ops.NormWeight
    conv1: ops.GemmWeight
    norm2: ops.NormWeight
    conv2: ops.GemmWeight
    shortcut: Optional[ops.GemmWeight]
    t_norm1: ops.NormWeight
    t_conv1: ops.GemmWeight
    t_norm2: ops.NormWeight
    t_conv2: ops.GemmWeight
    alpha: float                 # sigmoid(mix_factor)
    eps: float
    t_eps: float
    temb_off: int = -1           # column offset of the spatial time_emb_proj in the fused temb GEMM
    t_temb_off: int = -1
    cout: int = 0


@dataclass
class _Xf:
    norm: ops.NormWeight
    proj_in: ops.GemmWeight
    ln1: ops.NormWeight
    qkv: ops.GemmWeight
    to_out: ops.GemmWeight
    ln3: ops.NormWeight
    ff1: ops.GemmWeight
    ff2: ops.GemmWeight
    t_ln_in: ops.NormWeight
    t_ffin1: ops.GemmWeight
    t_ffin2: ops.GemmWeight
    t_ln1: ops.NormWeight
    t_qkv: ops.GemmWeight
    t_to_out: ops.GemmWeight
    t_ln3: ops.NormWeight
    t_ff1: ops.GemmWeight
    t_ff2: ops.GemmWeight
    proj_out: ops.GemmWeight
    alpha: float
    c: int
    heads: int
    # cross-attention (context length 1) collapsed to vectors
    x_to_v: ops.GemmWeight
    x_to_out: ops.GemmWeight
    tx_to_v: ops.GemmWeight
    tx_to_out: ops.GemmWeight
    pos_l1: ops.GemmWeight
    pos_l2: ops.GemmWeight
    pos: Optional[torch.Tensor] = None        # fp16 [F, C] frame-position embedding (model constant)
    xvec: Optional[torch.Tensor] = None       # fp16 [nb, C] spatial attn2 output per sample (request constant)
    txvec: Optional[torch.Tensor] = None      # fp16 [nb, C] temporal attn2 output per sample


class _Common:
    def __init__(self, device):
        self.device = torch.device(device)

    # ---- weight helpers
    def _lin(self, sd: SD, name: str, geglu: bool = False) -> ops.GemmWeight:
        return ops.prep_linear(sd[name + ".weight"], sd.get(name + ".bias"), self.device, geglu=geglu)

    def _conv(self, sd: SD, name: str) -> ops.GemmWeight:
        return ops.prep_conv(sd[name + ".weight"], sd.get(name + ".bias"), self.device)

    def _norm(self, sd: SD, name: str) -> ops.NormWeight:
        return ops.prep_norm(sd[name + ".weight"], sd[name + ".bias"], self.device)

    def _res(self, sd: SD, p: str, eps: float, t_eps: Optional[float] = None, switch: bool = False) -> _Res:
        s, t = p + ".spatial_res_block", p + ".temporal_res_block"
        alpha = float(torch.sigmoid(sd[p + ".time_mixer.mix_factor"].float())[0])
        if switch:
            alpha = 1.0 - alpha
        sc = self._conv(sd, s + ".conv_shortcut") if (s + ".conv_shortcut.weight") in sd else None
        conv2 = self._conv(sd, s + ".conv2")
        return _Res(self._norm(sd, s + ".norm1"), self._conv(sd, s + ".conv1"), self._norm(sd, s + ".norm2"), conv2, sc,
                    self._norm(sd, t + ".norm1"), self._conv(sd, t + ".conv1"), self._norm(sd, t + ".norm2"),
                    self._conv(sd, t + ".conv2"), alpha, eps, t_eps if t_eps is not None else eps, cout=conv2.n_real)

    # ---- SpatioTemporalResBlock
    def resblock(self, x: torch.Tensor, rb: _Res, nb: int, F: int, temb: Optional[torch.Tensor]) -> torch.Tensor:
        n, h, w, cin = x.shape
        S = h * w
        tab = tab_t = None
        if temb is not None:
            tab = temb[:, rb.temb_off: rb.temb_off + rb.cout]
            tab_t = temb[:, rb.t_temb_off: rb.t_temb_off + rb.cout]
        a = ops.group_norm(x, rb.norm1, n, rb.eps, True)
        hdn = ops.conv3x3(a, rb.conv1, tab=tab, tab_div=S, tab_mod=n)
        a = ops.group_norm(hdn, rb.norm2, n, rb.eps, True)
        sc = x if rb.shortcut is None else ops.conv1x1(x, rb.shortcut)
        xs = ops.conv3x3(a, rb.conv2, res=sc)                                  # spatial block output
        c = xs.shape[-1]
        a = ops.group_norm(xs, rb.t_norm1, nb, rb.t_eps, True)                 # stats over (C/32, F, h, w)
        hdn = ops.tconv3(a.view(nb, F, S, c), rb.t_conv1, tab=tab_t, tab_div=S, tab_mod=n)
        a = ops.group_norm(hdn, rb.t_norm2, nb, rb.t_eps, True)
        # alpha*x_s + (1-alpha)*(x_s + conv + b) = x_s + (1-alpha)*(conv + b)
        out = ops.tconv3(a.view(nb, F, S, c), rb.t_conv2, res=xs.view(nb, F, S, c), s_acc=1.0 - rb.alpha, s_res=1.0)
        return out.view(n, h, w, c)


    # ---- single-head (d = C) spatial attention with GroupNorm and residual (VAE mid blocks, unet_3d_blocks.py:958-969 /
    #      unet_2d_blocks.py UNetMidBlock2D): scores and P@V run on the implicit-GEMM kernel, one image at a time
    def single_head_attention(self, x: torch.Tensor, norm: ops.NormWeight, wq: ops.GemmWeight, wk: ops.GemmWeight,
                              wv: ops.GemmWeight, wo: ops.GemmWeight) -> torch.Tensor:
        n, h, w, c = x.shape
        S = h * w
        if S % 64:
            raise ValueError("VAE attention needs h*w % 64 == 0")
        g = ops.group_norm(x, norm, n, 1e-6, False).view(n * S, c)
        q, k, v = ops.linear(g, wq), ops.linear(g, wk), ops.linear(g, wv)
        vt = ops.transpose(v.view(n, S, c))                                   # [n, c, S]
        o = torch.empty(n * S, c, dtype=torch.float16, device=x.device)
        scores = torch.empty(S, S, dtype=torch.float16, device=x.device)
        for f in range(n):
            kf = ops.GemmWeight(w=k[f * S:(f + 1) * S], bias=None, n=S, k=c, cin=c, taps=1, bn=ops.pick_bn(S), n_real=S)
            ops.linear(q[f * S:(f + 1) * S], kf, out=scores)
            ops.softmax_rows_(scores, c ** -0.5)
            vf = ops.GemmWeight(w=vt[f], bias=None, n=c, k=S, cin=S, taps=1, bn=ops.pick_bn(c), n_real=c)
            ops.linear(scores, vf, out=o[f * S:(f + 1) * S])
        out = ops.linear(o, wo, res=x.view(n * S, c))
        return out.view(n, h, w, c)


# =========================================================================== UNet
class UNetEngine(_Common):
    """The spatio-temporal UNet of the SVD-dagger world model (micro_cond action strategy)."""

    def __init__(self, cfg: UNetConfig, sd: SD, device="cuda"):
        super().__init__(device)
        cfg.validate()
        self.cfg = cfg
        boc, n = cfg.block_out_channels, len(cfg.block_out_channels)
        self.conv_in = self._conv(sd, "conv_in")
        self.time_l1, self.time_l2 = self._lin(sd, "time_embedding.linear_1"), self._lin(sd, "time_embedding.linear_2")
        self.act_proj = self._lin(sd, "add_action_proj.proj")
        self.act_l1, self.act_l2 = self._lin(sd, "add_embedding_action.linear_1"), self._lin(sd, "add_embedding_action.linear_2")
        self.noise_l1, self.noise_l2 = self._lin(sd, "add_embedding_noise.linear_1"), self._lin(sd, "add_embedding_noise.linear_2")
        self.res: Dict[str, _Res] = {}
        self.xf: Dict[str, _Xf] = {}
        self.down_conv: Dict[int, ops.GemmWeight] = {}
        self.up_conv: Dict[int, ops.GemmWeight] = {}
        rheads = list(reversed(cfg.num_attention_heads))
        rev = list(reversed(boc))
        for i in range(n):
            for j in range(cfg.layers_per_block):
                self.res[f"down.{i}.{j}"] = self._res(sd, f"down_blocks.{i}.resnets.{j}", 1e-6 if i < n - 1 else 1e-5)
                if i < n - 1:
                    self.xf[f"down.{i}.{j}"] = self._xf(sd, f"down_blocks.{i}.attentions.{j}", boc[i], cfg.num_attention_heads[i])
            if i < n - 1:
                self.down_conv[i] = self._conv(sd, f"down_blocks.{i}.downsamplers.0.conv")
        self.res["mid.0"] = self._res(sd, "mid_block.resnets.0", 1e-5)
        self.xf["mid"] = self._xf(sd, "mid_block.attentions.0", boc[-1], cfg.num_attention_heads[-1])
        self.res["mid.1"] = self._res(sd, "mid_block.resnets.1", 1e-5)
        for i in range(n):
            for j in range(cfg.layers_per_block + 1):
                self.res[f"up.{i}.{j}"] = self._res(sd, f"up_blocks.{i}.resnets.{j}", 1e-6)
                if i > 0:
                    self.xf[f"up.{i}.{j}"] = self._xf(sd, f"up_blocks.{i}.attentions.{j}", rev[i], rheads[i])
            if i < n - 1:
                self.up_conv[i] = self._conv(sd, f"up_blocks.{i}.upsamplers.0.conv")
        self.norm_out = self._norm(sd, "conv_norm_out")
        self.conv_out = self._conv(sd, "conv_out")
        self.out_stride = self.conv_out.n                       # padded output channels (>= 4)
        # all time_emb_proj (spatial + temporal of every resblock) fused into one GEMM per step
        ws, bs, off = [], [], 0
        names = {"down": "down_blocks", "up": "up_blocks"}
        for key, rb in self.res.items():
            parts = key.split(".")
            p = ("mid_block.resnets." + parts[1]) if parts[0] == "mid" else f"{names[parts[0]]}.{parts[1]}.resnets.{parts[2]}"
            for sub, attr in (("spatial_res_block", "temb_off"), ("temporal_res_block", "t_temb_off")):
                w_, b_ = sd[f"{p}.{sub}.time_emb_proj.weight"], sd[f"{p}.{sub}.time_emb_proj.bias"]
                setattr(rb, attr, off)
                off += w_.shape[0]
                ws.append(w_.float()); bs.append(b_.float())
        pad = (-off) % 256
        if pad:
            ws.append(torch.zeros(pad, cfg.time_embed_dim, device=ws[0].device)); bs.append(torch.zeros(pad, device=ws[0].device))
        self.temb_all = ops.prep_linear(torch.cat(ws), torch.cat(bs), self.device)
        # frame-position embeddings are model constants: time_pos_embed(sincos(arange(F)))  (transformer_temporal.py:329-340)
        F = cfg.num_frames
        for xf in self.xf.values():
            t_emb = sinusoid(torch.arange(F), xf.c).to(torch.float16).to(self.device)
            xf.pos = ops.linear(ops.linear(t_emb, xf.pos_l1, act=1), xf.pos_l2)
        self._cond_ready = False
        self.nb = 0
        self._graph = None            # (cuda graph, static x_in, static t_emb, static v) of forward_graphed

    def _xf(self, sd: SD, p: str, c: int, heads: int) -> _Xf:
        b, t = p + ".transformer_blocks.0", p + ".temporal_transformer_blocks.0"

        def qkv(prefix):
            w = torch.cat([sd[f"{prefix}.to_q.weight"], sd[f"{prefix}.to_k.weight"], sd[f"{prefix}.to_v.weight"]])
            return ops.prep_linear(w, None, self.device)

        alpha = float(torch.sigmoid(sd[p + ".time_mixer.mix_factor"].float())[0])
        return _Xf(
            norm=self._norm(sd, p + ".norm"), proj_in=self._lin(sd, p + ".proj_in"),
            ln1=self._norm(sd, b + ".norm1"), qkv=qkv(b + ".attn1"), to_out=self._lin(sd, b + ".attn1.to_out.0"),
            ln3=self._norm(sd, b + ".norm3"), ff1=self._lin(sd, b + ".ff.net.0.proj", geglu=True), ff2=self._lin(sd, b + ".ff.net.2"),
            t_ln_in=self._norm(sd, t + ".norm_in"), t_ffin1=self._lin(sd, t + ".ff_in.net.0.proj", geglu=True),
            t_ffin2=self._lin(sd, t + ".ff_in.net.2"), t_ln1=self._norm(sd, t + ".norm1"), t_qkv=qkv(t + ".attn1"),
            t_to_out=self._lin(sd, t + ".attn1.to_out.0"), t_ln3=self._norm(sd, t + ".norm3"),
            t_ff1=self._lin(sd, t + ".ff.net.0.proj", geglu=True), t_ff2=self._lin(sd, t + ".ff.net.2"),
            proj_out=self._lin(sd, p + ".proj_out"), alpha=alpha, c=c, heads=heads,
            x_to_v=self._lin(sd, b + ".attn2.to_v"), x_to_out=self._lin(sd, b + ".attn2.to_out.0"),
            tx_to_v=self._lin(sd, t + ".attn2.to_v"), tx_to_out=self._lin(sd, t + ".attn2.to_out.0"),
            pos_l1=self._lin(sd, p + ".time_pos_embed.linear_1"), pos_l2=self._lin(sd, p + ".time_pos_embed.linear_2"))

    # ---- per-request constants -------------------------------------------------------------
    def set_condition(self, image_embedding: torch.Tensor, noise_aug_strength: float, action_ids: torch.Tensor,
                      cfg_rows: bool = True) -> None:
        """image_embedding [1 or ctx_len, D] fp (CLIP projection of the conditioning frame);
        action_ids [F, A] (cumulative action matrix of one branch, utils/svd_utils.py:594-632).
        Builds every step-invariant table: collapsed cross-attention vectors and the action /
        noise-augmentation parts of the time embedding (unet :467-487)."""
        cfg = self.cfg
        if image_embedding.dim() != 2 or image_embedding.shape[0] != 1:
            raise ValueError("cross-attention collapse needs a length-1 context (num_past_obs == 1)")
        dev = self.device
        nb_new = 2 if cfg_rows else 1
        same_nb = self._cond_ready and nb_new == self.nb
        if not same_nb:
            self._graph = None
        self.nb = nb_new
        F = cfg.num_frames
        kx = self.xf[next(iter(self.xf))].x_to_v.k
        ctx = torch.zeros(self.nb, kx, dtype=torch.float16, device=dev)     # CFG row 0 = zeros (pipeline :221-227)
        ctx[-1, : image_embedding.shape[1]] = image_embedding[0].to(dev, torch.float16)
        for xf in self.xf.values():
            # written in place when the buffers exist: a captured CUDA graph keeps pointing at them
            xf.xvec = ops.linear(ops.linear(ctx, xf.x_to_v), xf.x_to_out, out=xf.xvec if same_nb else None)
            xf.txvec = ops.linear(ops.linear(ctx, xf.tx_to_v), xf.tx_to_out, out=xf.txvec if same_nb else None)
        # action embedding (embeddings.py:912-945): cos/sin features of fp16 ids, element-major
        x = action_ids.to(torch.float16).float()
        feats = torch.stack([torch.cos(x), torch.sin(x), torch.cos(2.0 * x), torch.sin(2.0 * x), torch.cos(4.0 * x),
                             torch.sin(4.0 * x), torch.cos(x * 6.0), torch.sin(x * 6.0), torch.cos(x * 8.0),
                             torch.sin(x * 8.0), torch.cos(x * 10.0), torch.sin(x * 10.0)], dim=-1).reshape(F, -1)
        fpad = torch.zeros(F, self.act_proj.k, dtype=torch.float16, device=dev)
        fpad[:, : feats.shape[1]] = feats.to(torch.float16).to(dev)
        act = ops.linear(ops.linear(ops.linear(fpad, self.act_proj), self.act_l1, act=1), self.act_l2)      # [F, 1280]
        act_full = act.repeat(self.nb, 1).contiguous()                     # both CFG rows carry the same actions (unet :479-481)
        if same_nb:
            self.const_act.copy_(act_full)
        else:
            self.const_act = act_full
        nz = sinusoid(torch.tensor([noise_aug_strength], dtype=torch.float16), cfg.addition_time_embed_dim)
        nz = ops.linear(ops.linear(nz.to(torch.float16).to(dev), self.noise_l1, act=1), self.noise_l2)       # [1, 1280]
        nz_full = nz.repeat(self.nb * F, 1).contiguous()
        if same_nb:
            self.const_noise.copy_(nz_full)
        else:
            self.const_noise = nz_full
        self._cond_ready = True

    # ---- TransformerSpatioTemporalModel
    def transformer(self, x: torch.Tensor, xf: _Xf, nb: int, F: int) -> torch.Tensor:
        n, h, w, c = x.shape
        S, M = h * w, n * h * w
        g = ops.group_norm(x, xf.norm, n, 1e-6, False)
        hs = ops.linear(g.view(M, c), xf.proj_in)
        # spatial block: self-attention, (collapsed) cross-attention, GEGLU MLP
        qkv = ops.linear(ops.layer_norm(hs, xf.ln1), xf.qkv)
        att = ops.attention_spatial(qkv, n, S, c)
        hs = ops.linear(att, xf.to_out, res=hs, tab=xf.xvec, tab_div=F * S, tab_mod=nb)
        hs = ops.linear(ops.linear(ops.layer_norm(hs, xf.ln3), xf.ff1), xf.ff2, res=hs)
        # temporal block on h + frame-position embedding (rows stay in (b f s) order; kernels stride over f)
        u = ops.linear(ops.layer_norm(hs, xf.t_ln_in, tab=xf.pos, tab_div=S, tab_mod=F), xf.t_ffin1)
        hm = ops.linear(u, xf.t_ffin2, res=hs, tab=xf.pos, tab_div=S, tab_mod=F)
        qkv = ops.linear(ops.layer_norm(hm, xf.t_ln1), xf.t_qkv)
        att = ops.attention_temporal(qkv, nb, F, S, c)
        hm = ops.linear(att, xf.t_to_out, res=hm, tab=xf.txvec, tab_div=F * S, tab_mod=nb)
        u = ops.linear(ops.layer_norm(hm, xf.t_ln3), xf.t_ff1)
        # AlphaBlender: alpha*h_spatial + (1-alpha)*(ff + h_mix)
        hb = ops.linear(u, xf.t_ff2, res=hm, aux=hs, s_acc=1.0 - xf.alpha, s_res=1.0 - xf.alpha, s_aux=xf.alpha)
        out = ops.linear(hb, xf.proj_out, res=x.view(M, c))
        return out.view(n, h, w, c)

    # ---- full forward
    def forward(self, x_in: torch.Tensor, t_emb: torch.Tensor, taps: Optional[dict] = None) -> torch.Tensor:
        """x_in fp16 [nb*F, h, w, 64] (channels 0..7 real); t_emb fp16 [1, C0] sinusoid of the timestep.
        Returns v fp16 [nb*F, h, w, out_stride] (channels 0..3 real)."""
        if not self._cond_ready:
            raise RuntimeError("call set_condition() first")
        cfg, nb, F = self.cfg, self.nb, self.cfg.num_frames
        n_blk = len(cfg.block_out_channels)
        te = ops.linear(ops.linear(t_emb, self.time_l1, act=1), self.time_l2)                  # [1, 1280]
        emb_s = ops.emb_silu(te, self.const_act, self.const_noise)                               # silu(emb) [nb*F, 1280]
        temb = ops.linear(emb_s, self.temb_all)                                                  # every time_emb_proj at once
        x = ops.conv3x3(x_in, self.conv_in)
        skips = [x]
        for i in range(n_blk):
            for j in range(cfg.layers_per_block):
                x = self.resblock(x, self.res[f"down.{i}.{j}"], nb, F, temb)
                if i < n_blk - 1:
                    x = self.transformer(x, self.xf[f"down.{i}.{j}"], nb, F)
                skips.append(x)
                if taps is not None:
                    taps[f"down.{i}.{j}"] = x
            if i < n_blk - 1:
                x = ops.conv3x3_s2(x, self.down_conv[i])
                skips.append(x)
        x = self.resblock(x, self.res["mid.0"], nb, F, temb)
        x = self.transformer(x, self.xf["mid"], nb, F)
        x = self.resblock(x, self.res["mid.1"], nb, F, temb)
        if taps is not None:
            taps["mid"] = x
        for i in range(n_blk):
            for j in range(cfg.layers_per_block + 1):
                x = ops.concat_channels(x, skips.pop())
                x = self.resblock(x, self.res[f"up.{i}.{j}"], nb, F, temb)
                if i > 0:
                    x = self.transformer(x, self.xf[f"up.{i}.{j}"], nb, F)
                if taps is not None:
                    taps[f"up.{i}.{j}"] = x
            if i < n_blk - 1:
                x = ops.conv3x3(ops.upsample_nearest2x(x), self.up_conv[i])
        x = ops.group_norm(x, self.norm_out, x.shape[0], 1e-5, True)
        return ops.conv3x3(x, self.conv_out)


    def forward_graphed(self, x_in: torch.Tensor, t_emb: torch.Tensor) -> torch.Tensor:
        """Same as forward(), replayed from a CUDA graph captured on first use for this input shape: removes the
        ~1600 per-launch gaps of a step.  x_in / t_emb are copied into the graph's static inputs; the returned
        tensor is the graph's static output (valid until the next replay)."""
        g = self._graph
        if g is None or g[1].shape != x_in.shape:
            sx, st = x_in.clone(), t_emb.clone()
            self.forward(sx, st)                                   # warm-up outside capture (lazy inits, allocator)
            torch.cuda.synchronize()
            from . import lib as L
            graph = torch.cuda.CUDAGraph()
            n0 = L.COUNTERS["gpu_launches"]
            with torch.cuda.graph(graph):
                sv = self.forward(sx, st)
            g = self._graph = (graph, sx, st, sv, L.COUNTERS["gpu_launches"] - n0)
            L.COUNTERS["gpu_launches"] = n0                        # capture launched nothing
        g[1].copy_(x_in)
        g[2].copy_(t_emb)
        g[0].replay()
        from . import lib as L
        L.COUNTERS["gpu_launches"] += g[4]                         # kernels of this library inside the replayed graph
        return g[3]


# =========================================================================== temporal VAE decoder
class VAEDecoderEngine(_Common):
    """TemporalDecoder of AutoencoderKLTemporalDecoder; `sd` keys are relative to `decoder.`."""

    def __init__(self, cfg: VAEConfig, sd: SD, device="cuda"):
        super().__init__(device)
        cfg.validate()
        self.cfg = cfg
        boc = cfg.block_out_channels
        kw = dict(eps=1e-6, t_eps=1e-5, switch=True)
        self.conv_in = self._conv(sd, "conv_in")
        self.mid0 = self._res(sd, "mid_block.resnets.0", **kw)
        self.mid1 = self._res(sd, "mid_block.resnets.1", **kw)
        a = "mid_block.attentions.0"
        self.a_norm = self._norm(sd, a + ".group_norm")
        self.a_q, self.a_k, self.a_v = self._lin(sd, a + ".to_q"), self._lin(sd, a + ".to_k"), self._lin(sd, a + ".to_v")
        self.a_out = self._lin(sd, a + ".to_out.0")
        self.up: List[List[_Res]] = []
        self.up_conv: Dict[int, ops.GemmWeight] = {}
        for i in range(len(boc)):
            self.up.append([self._res(sd, f"up_blocks.{i}.resnets.{j}", **kw) for j in range(cfg.layers_per_block + 1)])
            if i < len(boc) - 1:
                self.up_conv[i] = self._conv(sd, f"up_blocks.{i}.upsamplers.0.conv")
        self.norm_out = self._norm(sd, "conv_norm_out")
        self.conv_out = self._conv(sd, "conv_out")
        self.tco_w = sd["time_conv_out.weight"].float().reshape(cfg.out_channels, cfg.out_channels, 3).contiguous().to(self.device)
        self.tco_b = sd["time_conv_out.bias"].float().contiguous().to(self.device)

    def attention(self, x: torch.Tensor) -> torch.Tensor:
        return self.single_head_attention(x, self.a_norm, self.a_q, self.a_k, self.a_v, self.a_out)

    def decode(self, z: torch.Tensor, num_frames: int, out_u8: Optional[torch.Tensor] = None,
               out_f32: Optional[torch.Tensor] = None) -> None:
        """z fp16 [F, h, w, 64] (channels 0..3 = latents / scaling_factor).  Writes uint8 [F,3,H,W]
        (postprocess_video quantisation) and / or fp32 [F,3,H,W] (decoder output before quantisation)."""
        from . import lib as L
        F = num_frames
        nb = z.shape[0] // F
        x = ops.conv3x3(z, self.conv_in)
        x = self.resblock(x, self.mid0, nb, F, None)
        x = self.attention(x)
        x = self.resblock(x, self.mid1, nb, F, None)
        for i, blocks in enumerate(self.up):
            for rb in blocks:
                x = self.resblock(x, rb, nb, F, None)
            if i in self.up_conv:
                x = ops.conv3x3(ops.upsample_nearest2x(x), self.up_conv[i])
        x = ops.group_norm(x, self.norm_out, x.shape[0], 1e-6, True)
        y = ops.conv3x3(x, self.conv_out)                                       # [F, H, W, 32] (3 real channels)
        n, H, W, cs = y.shape
        lib = L.require_device()
        for b in range(nb):
            yb = y[b * F:(b + 1) * F]
            L.check(lib.wiw_vae_finish(yb.data_ptr(), self.tco_w.data_ptr(), self.tco_b.data_ptr(),
                                       out_u8[b * F:(b + 1) * F].data_ptr() if out_u8 is not None else None,
                                       out_f32[b * F:(b + 1) * F].data_ptr() if out_f32 is not None else None,
                                       F, H * W, cs, L.stream_ptr()))


# =========================================================================== SD-VAE encoder (conditioning frame)
class VAEEncoderEngine(_Common):
    """`AutoencoderKLTemporalDecoder.encode(x).latent_dist.mode()` on the sm_100a kernels: the SD VAE Encoder
    (DP/models/autoencoders/vae.py:47-182; DownEncoderBlock2D / UNetMidBlock2D in DP/models/unets/unet_2d_blocks.py) followed by
    quant_conv (autoencoder_kl_temporal_decoder.py:316-341), posterior mean only.  `sd` holds `encoder.*` and `quant_conv.*`.

    OPT-IN (`wm_worker --own_vae_encoder`): the reference runs this stage in fp32 (`force_upcast`), here activations are fp16
    with fp32 accumulation like the rest of the path.  Host logic is checked on the CPU against the reference golden
    (tests/test_engine_host_cpu.py); it is not the default until it has been measured on a real checkpoint."""

    def __init__(self, cfg: VAEConfig, sd: SD, device="cuda"):
        super().__init__(device)
        cfg.validate()
        self.cfg = cfg
        boc = cfg.block_out_channels
        self.conv_in = self._conv(sd, "encoder.conv_in")
        self.down: List[List[dict]] = []
        self.down_conv: Dict[int, ops.GemmWeight] = {}
        for i in range(len(boc)):
            self.down.append([self._res2d(sd, f"encoder.down_blocks.{i}.resnets.{j}") for j in range(cfg.layers_per_block)])
            if i < len(boc) - 1:
                self.down_conv[i] = self._conv(sd, f"encoder.down_blocks.{i}.downsamplers.0.conv")
        self.mid0 = self._res2d(sd, "encoder.mid_block.resnets.0")
        self.mid1 = self._res2d(sd, "encoder.mid_block.resnets.1")
        a = "encoder.mid_block.attentions.0"
        self.a_norm = self._norm(sd, a + ".group_norm")
        self.a_q, self.a_k, self.a_v = self._lin(sd, a + ".to_q"), self._lin(sd, a + ".to_k"), self._lin(sd, a + ".to_v")
        self.a_out = self._lin(sd, a + ".to_out.0")
        self.norm_out = self._norm(sd, "encoder.conv_norm_out")
        # quant_conv (1x1, 2L -> 2L) folded into conv_out in fp32 at load time: Wq @ W_out, Wq @ b_out + bq (exact algebra, one launch less)
        wq = sd["quant_conv.weight"].float().flatten(1)                                        # [2L, 2L]
        w_out, b_out = sd["encoder.conv_out.weight"].float(), sd["encoder.conv_out.bias"].float()
        w_fused = torch.einsum("om,mikl->oikl", wq, w_out)
        b_fused = wq @ b_out + sd["quant_conv.bias"].float()
        self.conv_out = ops.prep_conv(w_fused, b_fused, self.device)
        self.latent_channels = cfg.latent_channels

    def _res2d(self, sd: SD, p: str) -> dict:
        sc = self._conv(sd, p + ".conv_shortcut") if (p + ".conv_shortcut.weight") in sd else None
        return dict(norm1=self._norm(sd, p + ".norm1"), conv1=self._conv(sd, p + ".conv1"), norm2=self._norm(sd, p + ".norm2"),
                    conv2=self._conv(sd, p + ".conv2"), shortcut=sc)

    def _resblock2d(self, x: torch.Tensor, r: dict) -> torch.Tensor:
        n = x.shape[0]
        a = ops.group_norm(x, r["norm1"], n, 1e-6, True)
        hdn = ops.conv3x3(a, r["conv1"])
        a = ops.group_norm(hdn, r["norm2"], n, 1e-6, True)
        sc = x if r["shortcut"] is None else ops.conv1x1(x, r["shortcut"])
        return ops.conv3x3(a, r["conv2"], res=sc)

    @torch.no_grad()
    def encode_mode(self, image: torch.Tensor) -> torch.Tensor:
        """image float [B, 3, H, W] in [-1, 1] (H, W multiples of 8) -> posterior mean fp16 [B, L, H/8, W/8] (not scaled)."""
        x = ops.nchw_to_nhwc(image.to(self.device, torch.float16).contiguous(), 64)            # 3 real channels, zero-padded to 64
        x = ops.conv3x3(x, self.conv_in)
        for i, blocks in enumerate(self.down):
            for r in blocks:
                x = self._resblock2d(x, r)
            if i in self.down_conv:
                x = ops.conv3x3_s2(x, self.down_conv[i], asym=True)
        x = self._resblock2d(x, self.mid0)
        x = self.single_head_attention(x, self.a_norm, self.a_q, self.a_k, self.a_v, self.a_out)
        x = self._resblock2d(x, self.mid1)
        x = ops.group_norm(x, self.norm_out, x.shape[0], 1e-6, True)
        moments = ops.conv3x3(x, self.conv_out)                                                 # [B, h, w, 32] (2L real channels)
        return ops.nhwc_to_nchw(moments, self.latent_channels)
