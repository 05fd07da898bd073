"""Host-side op wrappers: build the C-ABI descriptors for the sm_100a kernels.

Activations are fp16 channels-last torch CUDA tensors; every function launches on the
current torch stream and returns immediately.  No op here has a torch fallback.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import lib as L


# ----------------------------------------------------------------------------- weights
@dataclass
class GemmWeight:
    """fp16 [n, k] K-major weight (+ fp32 bias) ready for wiw_conv_gemm."""
    w: torch.Tensor
    bias: Optional[torch.Tensor]
    n: int
    k: int
    cin: int                 # channels per tap (multiple of 64, after padding)
    taps: int
    geglu: bool = False
    bn: int = 0
    n_real: int = 0          # un-padded output channels
    k_real: int = 0          # un-padded reduction length (taps * real input channels)


def pick_bn(n: int, geglu: bool = False) -> int:
    """N-tile width.  Measured on B200 (profiles/r01_mma_cost.md): an SS-mode tcgen05.mma with M=128 costs the same
    ~128 cycles for every N <= 256, so the cheapest tiling is the one with the FEWEST N-tiles; among those the
    narrowest width (least padding, most pipeline stages).  The last tile may be ragged (the kernel zero-fills and
    skips its padded columns); GEGLU tiles must divide N because value/gate columns are interleaved per tile."""
    step = 64 if geglu else 32
    if n % step:
        raise ValueError(f"N={n} is not a multiple of {step}")
    if geglu:
        for bn in range(256, 0, -step):
            if n % bn == 0:
                return bn
    tiles = -(-n // 256)
    for bn in range(step, 257, step):
        if -(-n // bn) == tiles:
            return bn
    raise AssertionError


def _pad_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def prep_linear(weight: torch.Tensor, bias: Optional[torch.Tensor], device, geglu: bool = False) -> GemmWeight:
    """nn.Linear weight [n, k] -> K-major fp16.  For GEGLU (`net.0.proj`, [8C, C]: first half value,
    second half gate -- activations.py:120) rows are interleaved per N-tile as [bn/2 value | bn/2 gate]."""
    n, k = weight.shape
    kp, npad = _pad_to(k, 64), _pad_to(n, 64 if geglu else 32)
    wd = weight.device                      # prepare where the checkpoint lives (GPU for the bench, CPU for tests)
    w = torch.zeros(npad, kp, dtype=torch.float32, device=wd)
    w[:n, :k] = weight.float()
    b = torch.zeros(npad, dtype=torch.float32, device=wd)
    if bias is not None:
        b[:n] = bias.float()
    bn = pick_bn(npad, geglu)
    if geglu:
        assert npad == n and n % 2 == 0
        half, hb = n // 2, bn // 2
        idx = []
        for t in range(n // bn):
            idx += list(range(t * hb, (t + 1) * hb)) + list(range(half + t * hb, half + (t + 1) * hb))
        idx = torch.tensor(idx, device=wd)
        w, b = w[idx], b[idx]
    return GemmWeight(w=w.to(device=device, dtype=torch.float16).contiguous(),
                      bias=(b.to(device) if bias is not None else None), n=npad, k=kp, cin=kp, taps=1,
                      geglu=geglu, bn=bn, n_real=n, k_real=k)


def prep_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], device) -> GemmWeight:
    """Conv2d [cout, cin, 3, 3] / [cout, cin, 1, 1] or Conv3d [cout, cin, 3, 1, 1] -> [cout, taps*cin_p]
    with tap-major, channel-minor K (tap index = ky*3 + kx, or kt)."""
    cout, cin = weight.shape[:2]
    taps = int(torch.tensor(weight.shape[2:]).prod())
    cp, npad = _pad_to(cin, 64), _pad_to(cout, 32)
    wd = weight.device
    w = torch.zeros(npad, taps, cp, dtype=torch.float32, device=wd)
    w[:cout, :, :cin] = weight.float().reshape(cout, cin, taps).permute(0, 2, 1)
    b = torch.zeros(npad, dtype=torch.float32, device=wd)
    if bias is not None:
        b[:cout] = bias.float()
    return GemmWeight(w=w.reshape(npad, taps * cp).to(device=device, dtype=torch.float16).contiguous(),
                      bias=b.to(device), n=npad, k=taps * cp, cin=cp, taps=taps, bn=pick_bn(npad), n_real=cout, k_real=taps * cin)


# ----------------------------------------------------------------------------- GEMM launches
PROFILE = None        # set to a list to record (algorithmic flops, start event, end event) per conv_gemm launch


_SMART_BOXES = os.environ.get("WIW_BOX3", "1") != "0"      # development aid for A/B runs: 0 = bx x by x 1 tiles only
_WAVE_BN = os.environ.get("WIW_WAVE_BN", "1") != "0"       # development aid for A/B runs: 0 = the weight's own N-tile width always
_FORCE_BN = int(os.environ.get("WIW_FORCE_BN", "0"))
_SLOTS = 148                                                # SMs of a B200; CTA pairs = half


def launch_bn(gw: "GemmWeight", m_tiles: int) -> int:
    """N-tile width of ONE launch, chosen against wave quantisation.  The kernel is persistent with one CTA (pair) per SM
    (pair): a launch takes ceil(tiles / slots) waves of tiles whose duration grows with bn (a cta_group::2 MMA costs bn / 2 cycles
    for bn >= 128, profiles/r02_mma_cost_v2.txt), so the narrowest width that removes a nearly empty last wave wins:
    the 9x16 level (35 M-tiles) runs N = 1280 as 90 pair-tiles of 256 = 2 waves on 74 pairs, but as 144 pair-tiles of 160 in the
    same 2 waves at 62 % of the time; the 18x32 level goes from 5 waves of 256 to 7 of 160 (-12 %).  GEGLU weights are interleaved
    per tile at load time and keep their width; the last N-tile may be ragged (its padded columns still cost a full tile)."""
    if not _WAVE_BN or gw.geglu or gw.n < 128:
        return gw.bn
    if _FORCE_BN:                                                                      # development aid: measure one width everywhere
        return min(_FORCE_BN, gw.n)
    def cost(bn: int):
        n_tiles = -(-gw.n // bn)
        total = m_tiles * n_tiles
        paired = total >= 2 * _SLOTS and not (gw.k <= 320 and gw.n <= 320)          # the library's cta_group::2 rule
        units, slots = ((-(-m_tiles // 2)) * n_tiles, _SLOTS // 2) if paired else (total, _SLOTS)
        waves = -(-units // slots)
        return waves, waves * (bn + 24)                                                # + ~24 columns' worth of per-tile fixed cost

    waves0, cost0 = cost(gw.bn)
    if waves0 > 8:                                                                     # many waves: the last one hardly matters
        return gw.bn
    best, best_cost = gw.bn, cost0
    for bn in (128, 160, 192, 224, 256):
        if bn > gw.n:
            continue
        c = cost(bn)[1]
        if c < 0.92 * cost0 and (c < best_cost or (c == best_cost and bn > best)):    # switch only for a clear (>= 8 %) predicted gain
            best, best_cost = bn, c
    return best


def _box(width: int) -> Tuple[int, int]:
    """M-tile = by rows of bx pixels (bx*by = 128); bx = largest power of two <= min(width, 128)."""
    bx = 1
    while bx * 2 <= min(width, 128):
        bx *= 2
    return bx, 128 // bx


def _box3(w: int, h: int, n: int) -> Tuple[int, int, int]:
    """M-tile of a spatial conv = bx pixels x by rows x bz images (product 128), bx = widest power of two <= min(w, 128):
    the (by, bz) split that wastes the fewest tile rows -- a 9-row map fills 16 x 8 x 1 tiles to 56 %, 16 x 2 x 4 to 90 %;
    an 18-row map 32 x 4 x 1 to 90 %, 32 x 2 x 2 to 100 %.  Ties go to fewer images per tile."""
    bx, _ = _box(w)
    if not _SMART_BOXES:
        return bx, 128 // bx, 1
    best, best_u = (bx, 128 // bx, 1), -1.0
    by = 128 // bx
    bz = 1
    while by >= 1:
        u = (h / (-(-h // by) * by)) * (n / (-(-n // bz) * bz))
        if u > best_u + 1e-9:
            best, best_u = (bx, by, bz), u
        by //= 2
        bz *= 2
    return best


def _tbox(s: int, f: int) -> Tuple[int, int]:
    """M-tile of the temporal conv = bx pixels x by frames (bx*by = 128): the split that wastes the fewest tile rows
    (S = 144 at the lowest UNet level: 128 x 1 tiles are 56 % full, 16 x 8 tiles 87.5 %); ties go to the wider bx."""
    best, best_u = (128, 1), -1.0
    if not _SMART_BOXES:
        return best
    for bx in (128, 64, 32, 16, 8):
        by = 128 // bx
        u = (s / (-(-s // bx) * bx)) * (f / (-(-f // by) * by))
        if u > best_u + 1e-9:
            best, best_u = (bx, by), u
    return best


def _launch(a: torch.Tensor, a_dim, a_stride, gw: GemmWeight, out: torch.Tensor, out_dim, out_stride,
            taps: Sequence[Tuple[int, int, int, int]], bx: int, by: int,
            res: Optional[torch.Tensor], aux: Optional[torch.Tensor], tab: Optional[torch.Tensor],
            tab_div: int, tab_mod: int, s_acc: float, s_res: float, s_aux: float, use_bias: bool = True,
            act: int = 0, bz: int = 1) -> None:
    lib = L.require_device()
    d = L.GemmDesc()
    d.a = a.data_ptr()
    for i in range(5):
        d.a_dim[i] = int(a_dim[i]); d.a_stride[i] = int(a_stride[i])
    d.w = gw.w.data_ptr(); d.k_total = gw.k; d.n_total = gw.n
    d.out = out.data_ptr()
    for i in range(4):
        d.out_dim[i] = int(out_dim[i]); d.out_stride[i] = int(out_stride[i])
        d.res_stride[i] = int(out_stride[i]); d.aux_stride[i] = int(out_stride[i])
    d.res = res.data_ptr() if res is not None else None
    d.aux = aux.data_ptr() if aux is not None else None
    d.bias = gw.bias.data_ptr() if (gw.bias is not None and use_bias) else None
    d.tab = tab.data_ptr() if tab is not None else None
    d.tab_div, d.tab_mod = int(tab_div), int(tab_mod)
    if tab is not None:
        if tab.dtype != torch.float16 or tab.stride(-1) != 1:
            raise L.WiwError("tab must be fp16 with contiguous columns")
        d.tab_stride = int(tab.stride(0)) if tab.dim() == 2 else int(tab.shape[-1])
    d.act = int(act)
    d.s_acc, d.s_res, d.s_aux = float(s_acc), float(s_res), float(s_aux)
    m_tiles = (-(-int(out_dim[1]) // bx)) * (-(-int(out_dim[2]) // by)) * (-(-int(out_dim[3]) // bz))
    d.bx, d.by, d.bn, d.geglu = bx, by, launch_bn(gw, m_tiles), int(gw.geglu)
    d.bz = bz
    d.num_taps, d.chunks_per_tap = len(taps), gw.cin // 64
    for i, (cb, ox, p, oy) in enumerate(taps):
        d.tap_cbase[i], d.tap_ox[i], d.tap_p[i], d.tap_oy[i] = cb, ox, p, oy
    if PROFILE is None:
        L.check(lib.wiw_conv_gemm(d, L.stream_ptr()))
        return
    # bench.py roofline pass: CUDA events around every launch of the dominant kernel (on its own stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(lib.wiw_conv_gemm(d, L.stream_ptr()))
    e1.record()
    rows = int(out_dim[1]) * int(out_dim[2]) * int(out_dim[3])
    PROFILE.append((2.0 * rows * (gw.n_real or gw.n) * (gw.k_real or gw.k), e0, e1,
                    f"M={rows} K={gw.k} N={gw.n} taps={len(taps)} bn={d.bn}{' geglu' if gw.geglu else ''}"
                    f"{' res' if res is not None else ''}{' aux' if aux is not None else ''}{' tab' if tab is not None else ''}"))


def _chk(x: torch.Tensor) -> None:
    if x.dtype != torch.float16 or not x.is_cuda or not x.is_contiguous():
        raise L.WiwError(f"expected a contiguous fp16 CUDA tensor, got {x.dtype} {x.device} contiguous={x.is_contiguous()}")


def linear(x: torch.Tensor, gw: GemmWeight, res=None, aux=None, tab=None, tab_div=1, tab_mod=1,
           s_acc=1.0, s_res=1.0, s_aux=1.0, out: Optional[torch.Tensor] = None, act: int = 0) -> torch.Tensor:
    """y[M, n] = epilogue(x[M, k] @ W^T): nn.Linear (+ fused bias / residual / GEGLU / row-group table)."""
    _chk(x)
    m, k = x.shape
    assert k == gw.k, (k, gw.k)
    n_out = gw.n // 2 if gw.geglu else gw.n
    if out is None:
        out = torch.empty(m, n_out, dtype=torch.float16, device=x.device)
    _launch(x, (k, m, 1, 1, 1), (1, k, m * k, m * k, m * k), gw, out, (n_out, m, 1, 1), (1, n_out, m * n_out, m * n_out),
            [(0, 0, 0, 0)], 128, 1, res, aux, tab, tab_div, tab_mod, s_acc, s_res, s_aux, act=act)
    return out


def conv3x3(x: torch.Tensor, gw: GemmWeight, res=None, tab=None, tab_div=1, tab_mod=1,
            s_acc=1.0, s_res=1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv2d 3x3 stride 1 pad 1 on channels-last x[N, H, W, C] (C % 64 == 0)."""
    _chk(x)
    n, h, w, c = x.shape
    assert c == gw.cin and gw.taps == 9
    if out is None:
        out = torch.empty(n, h, w, gw.n, dtype=torch.float16, device=x.device)
    bx, by, bz = _box3(w, h, n)
    taps = [(0, kx - 1, 0, ky - 1) for ky in range(3) for kx in range(3)]
    _launch(x, (c, w, 1, h, n), (1, c, w * c, w * c, h * w * c), gw, out, (gw.n, w, h, n), (1, gw.n, w * gw.n, h * w * gw.n),
            taps, bx, by, res, None, tab, tab_div, tab_mod, s_acc, s_res, 1.0, bz=bz)
    return out


def conv1x1(x: torch.Tensor, gw: GemmWeight, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv2d 1x1 (conv_shortcut, resnet.py:311) == linear over pixels."""
    n, h, w, c = x.shape
    y = linear(x.view(n * h * w, c), gw, out=None if out is None else out.view(n * h * w, gw.n))
    return y.view(n, h, w, gw.n)


def conv3x3_s2(x: torch.Tensor, gw: GemmWeight, out: Optional[torch.Tensor] = None, asym: bool = False) -> torch.Tensor:
    """Conv2d 3x3 stride 2 pad 1 (Downsample2D, downsampling.py:69-146) through a space-to-depth view:
    input pixel (2Y+py, 2X+px) -> coords (c + px*C, X, py, Y, n); tap d=-1 reads parity 1 at X-1 / Y-1.
    asym=True: the VAE encoder's variant (downsampling.py `padding == 0` branch: F.pad(x, (0,1,0,1)) then stride 2 without
    padding), i.e. taps at rows 2Y, 2Y+1, 2Y+2 -- the row / column past the image is the tensor map's zero fill."""
    _chk(x)
    n, h, w, c = x.shape
    assert c == gw.cin and gw.taps == 9 and h % 2 == 0 and w % 2 == 0
    ho, wo = h // 2, w // 2
    if out is None:
        out = torch.empty(n, ho, wo, gw.n, dtype=torch.float16, device=x.device)
    bx, by, bz = _box3(wo, ho, n)

    def par(d):           # filter tap d in {0,1,2} -> (parity, coarse offset); symmetric padding: rows 2Y-1, 2Y, 2Y+1
        return [(0, 0), (1, 0), (0, 1)][d] if asym else [(1, -1), (0, 0), (1, 0)][d]

    taps = []
    for ky in range(3):
        py, oy = par(ky)
        for kx in range(3):
            px, ox = par(kx)
            taps.append((px * c, ox, py, oy))
    _launch(x, (2 * c, wo, 2, ho, n), (1, 2 * c, w * c, 2 * w * c, h * w * c), gw, out,
            (gw.n, wo, ho, n), (1, gw.n, wo * gw.n, ho * wo * gw.n), taps, bx, by,
            None, None, None, 1, 1, 1.0, 1.0, 1.0, bz=bz)
    return out


def tconv3(x: torch.Tensor, gw: GemmWeight, res=None, tab=None, tab_div=1, tab_mod=1,
           s_acc=1.0, s_res=1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Conv3d (3,1,1) pad (1,0,0) along frames on x[B, F, S, C] (TemporalResnetBlock, resnet.py:570,586)."""
    _chk(x)
    b, f, s, c = x.shape
    assert c == gw.cin and gw.taps == 3
    if out is None:
        out = torch.empty(b, f, s, gw.n, dtype=torch.float16, device=x.device)
    taps = [(0, 0, 0, kt - 1) for kt in range(3)]
    bx, by = _tbox(s, f)
    _launch(x, (c, s, 1, f, b), (1, c, s * c, s * c, f * s * c), gw, out, (gw.n, s, f, b), (1, gw.n, s * gw.n, f * s * gw.n),
            taps, bx, by, res, None, tab, tab_div, tab_mod, s_acc, s_res, 1.0)
    return out


# ----------------------------------------------------------------------------- norms
@dataclass
class NormWeight:
    gamma: torch.Tensor      # fp32 [C]
    beta: torch.Tensor       # fp32 [C]


def prep_norm(weight: torch.Tensor, bias: torch.Tensor, device) -> NormWeight:
    return NormWeight(weight.float().to(device).contiguous(), bias.float().to(device).contiguous())


_gn_ws = {}


def group_norm(x: torch.Tensor, nw: NormWeight, samples: int, eps: float, silu: bool, groups: int = 32,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.GroupNorm(32)(+SiLU) over x viewed as [samples][rows][C] (channels-last)."""
    _chk(x)
    lib = L.require_device()
    c = x.shape[-1]
    rows = x.numel() // c // samples
    if out is None:
        out = torch.empty_like(x)
    need = lib.wiw_groupnorm_workspace_floats(samples, groups)
    key = (x.device.index, torch.cuda.current_stream().cuda_stream)
    ws = _gn_ws.get(key)
    if ws is None or ws.numel() < need:
        ws = torch.empty(max(need, 1 << 16), dtype=torch.float32, device=x.device)
        _gn_ws[key] = ws
    L.check(lib.wiw_groupnorm(x.data_ptr(), out.data_ptr(), samples, rows, c, groups, nw.gamma.data_ptr(),
                              nw.beta.data_ptr(), float(eps), int(silu), ws.data_ptr(), L.stream_ptr()))
    return out


def layer_norm(x: torch.Tensor, nw: NormWeight, eps: float = 1e-5, tab: Optional[torch.Tensor] = None,
               tab_div: int = 1, tab_mod: int = 1, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """nn.LayerNorm over the last axis of x[M, C]; optional fp16 pre-add of tab[(row // tab_div) % tab_mod]."""
    _chk(x)
    lib = L.require_device()
    m, c = x.shape
    if out is None:
        out = torch.empty_like(x)
    L.check(lib.wiw_layernorm(x.data_ptr(), out.data_ptr(), m, c, nw.gamma.data_ptr(), nw.beta.data_ptr(), float(eps),
                              tab.data_ptr() if tab is not None else None, tab_div, tab_mod, L.stream_ptr()))
    return out


# ----------------------------------------------------------------------------- attention
def attention_spatial(qkv: torch.Tensor, images: int, s: int, c: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [images*S, 3C] -> [images*S, C]: softmax(q k^T / 8) v per (image, head)."""
    _chk(qkv)
    lib = L.require_device()
    assert qkv.shape == (images * s, 3 * c)
    if out is None:
        out = torch.empty(images * s, c, dtype=torch.float16, device=qkv.device)
    L.check(lib.wiw_attention_spatial(qkv.data_ptr(), out.data_ptr(), images, s, c, L.stream_ptr()))
    return out


def attention_temporal(qkv: torch.Tensor, b: int, f: int, s: int, c: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """qkv [B*F*S, 3C] (rows ordered b, f, s) -> [B*F*S, C]: attention over the F frames of each pixel."""
    _chk(qkv)
    lib = L.require_device()
    assert qkv.shape == (b * f * s, 3 * c)
    if out is None:
        out = torch.empty(b * f * s, c, dtype=torch.float16, device=qkv.device)
    L.check(lib.wiw_attention_temporal(qkv.data_ptr(), out.data_ptr(), b, f, s, c, L.stream_ptr()))
    return out


# ----------------------------------------------------------------------------- small glue
def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _chk(a); _chk(b)
    lib = L.require_device()
    c1, c2 = a.shape[-1], b.shape[-1]
    m = a.numel() // c1
    out = torch.empty(*a.shape[:-1], c1 + c2, dtype=torch.float16, device=a.device)
    L.check(lib.wiw_concat_channels(a.data_ptr(), b.data_ptr(), out.data_ptr(), m, c1, c2, L.stream_ptr()))
    return out


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    _chk(x)
    lib = L.require_device()
    n, h, w, c = x.shape
    out = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.float16, device=x.device)
    L.check(lib.wiw_upsample_nearest2x(x.data_ptr(), out.data_ptr(), n, h, w, c, L.stream_ptr()))
    return out


def emb_silu(te: torch.Tensor, act: torch.Tensor, noise: torch.Tensor) -> torch.Tensor:
    lib = L.require_device()
    r, c = act.shape
    out = torch.empty(r, c, dtype=torch.float16, device=act.device)
    L.check(lib.wiw_emb_silu(te.data_ptr(), act.data_ptr(), noise.data_ptr(), out.data_ptr(), r, c, L.stream_ptr()))
    return out


def softmax_rows_(s: torch.Tensor, scale: float) -> torch.Tensor:
    _chk(s)
    lib = L.require_device()
    rows, n = s.shape
    L.check(lib.wiw_softmax_rows(s.data_ptr(), rows, n, float(scale), L.stream_ptr()))
    return s


def transpose(x: torch.Tensor) -> torch.Tensor:
    """[batch, R, C] -> [batch, C, R]"""
    _chk(x)
    lib = L.require_device()
    b, r, c = x.shape
    out = torch.empty(b, c, r, dtype=torch.float16, device=x.device)
    L.check(lib.wiw_transpose(x.data_ptr(), out.data_ptr(), b, r, c, L.stream_ptr()))
    return out


def nchw_to_nhwc(x: torch.Tensor, cs: int) -> torch.Tensor:
    """fp16 [N, C, *spatial] -> channels-last [N, *spatial, cs] (zero-padded channels)."""
    _chk(x)
    lib = L.require_device()
    n, c = x.shape[:2]
    p = x.numel() // (n * c)
    out = torch.empty(n, *x.shape[2:], cs, dtype=torch.float16, device=x.device)
    L.check(lib.wiw_nchw_to_nhwc(x.data_ptr(), out.data_ptr(), n, c, p, cs, L.stream_ptr()))
    return out


def nhwc_to_nchw(x: torch.Tensor, c: int) -> torch.Tensor:
    """channels-last fp16 [N, *spatial, cs] -> [N, c, *spatial] (first c channels)."""
    _chk(x)
    lib = L.require_device()
    n, cs = x.shape[0], x.shape[-1]
    p = x.numel() // (n * cs)
    out = torch.empty(n, c, *x.shape[1:-1], dtype=torch.float16, device=x.device)
    L.check(lib.wiw_nhwc_to_nchw(x.data_ptr(), out.data_ptr(), n, c, p, cs, L.stream_ptr()))
    return out
