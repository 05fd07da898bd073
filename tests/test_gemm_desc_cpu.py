"""Host logic of the implicit-GEMM wrappers, checked WITHOUT a GPU.

`ops.linear / conv3x3 / conv3x3_s2 / tconv3` only build a `wiw_gemm_desc` (views, strides, filter-tap tables, weight repacking,
GEGLU interleave, epilogue operands) and hand it to the C ABI.  Here the C-ABI call is replaced by a numpy emulator that
executes the *documented meaning* of the descriptor (include/wiw.h) on host memory -- test infrastructure, never the
product -- and the result is compared with torch's own conv / linear.  What this pins on the CPU: the tap tables (zero
padding, the space-to-depth view of the stride-2 conv, the frame taps of the temporal conv), weight layouts, bias / table /
residual / aux / AlphaBlender semantics and the tile-shape choices staying inside the descriptor contract."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wiw_b200  # noqa: F401
from wiw_b200 import lib as L
from wiw_b200 import ops


def _arr(ptr, n, dtype):
    if not ptr:
        return None
    buf = (ctypes.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype)


class FakeLib:
    """Executes a wiw_gemm_desc the way include/wiw.h documents it (fp32 accumulation, fp16 output)."""

    def __init__(self):
        self.calls = []

    def wiw_conv_gemm(self, d, stream):
        ad, ast = list(d.a_dim), list(d.a_stride)
        od, ost = list(d.out_dim), list(d.out_stride)
        taps, cpt = d.num_taps, d.chunks_per_tap
        cin = cpt * 64
        assert d.k_total == taps * cin and d.bx * d.by * max(d.bz, 1) == 128 and d.bn % 32 == 0 and 32 <= d.bn <= 256
        a_span = 1 + sum((ad[i] - 1) * ast[i] for i in range(5))
        A = _arr(d.a, a_span, np.float16).astype(np.float32)
        Wm = _arr(d.w, d.n_total * d.k_total, np.float16).astype(np.float32).reshape(d.n_total, d.k_total)
        n_out = d.n_total // 2 if d.geglu else d.n_total
        assert od[0] == n_out
        X, Y, N = od[1], od[2], od[3]
        acc = np.zeros((N, Y, X, d.n_total), np.float32)
        xs, ys, ns = np.arange(X), np.arange(Y), np.arange(N)
        for t in range(taps):
            cb, ox, p, oy = d.tap_cbase[t], d.tap_ox[t], d.tap_p[t], d.tap_oy[t]
            xi, yi = xs + ox, ys + oy
            ok = ((xi >= 0) & (xi < ad[1]))[None, None, :] & ((yi >= 0) & (yi < ad[3]))[None, :, None] & np.ones((N, 1, 1), bool)
            base = (ns[:, None, None] * ast[4] + np.clip(yi, 0, ad[3] - 1)[None, :, None] * ast[3] + p * ast[2]
                    + np.clip(xi, 0, ad[1] - 1)[None, None, :] * ast[1])
            ch = cb + np.arange(cin)
            chok = ch < ad[0]                                                  # channels past the tensor read as zero (TMA OOB fill)
            patch = A[(base[..., None] + np.clip(ch, 0, ad[0] - 1) * ast[0])] * chok * ok[..., None]
            acc += patch @ Wm[:, t * cin:(t + 1) * cin].T
        if d.bias:
            acc += _arr(d.bias, d.n_total, np.float32)
        if d.tab:
            grow = (ns[:, None, None] * Y + ys[None, :, None]) * X + xs[None, None, :]
            rows = (grow // d.tab_div) % d.tab_mod
            stride = d.tab_stride if d.tab_stride > 0 else d.n_total
            tab = _arr(d.tab, (d.tab_mod - 1) * stride + d.n_total, np.float16).astype(np.float32)
            acc += tab[rows[..., None] * stride + np.arange(d.n_total)]
        if d.act == 1:
            acc = acc / (1.0 + np.exp(-acc))
        if d.geglu:
            bn, hb = d.bn, d.bn // 2
            v = acc.reshape(N, Y, X, d.n_total // bn, bn)
            gate = torch.from_numpy(np.ascontiguousarray(v[..., hb:]))
            acc = (v[..., :hb] * F.gelu(gate).numpy()).reshape(N, Y, X, n_out)
        out = acc * d.s_acc
        o_span = 1 + sum((od[i] - 1) * ost[i] for i in range(4))
        idx = (ns[:, None, None, None] * ost[3] + ys[None, :, None, None] * ost[2] + xs[None, None, :, None] * ost[1]
               + np.arange(n_out)[None, None, None, :] * ost[0])
        if d.res:
            out = out + d.s_res * _arr(d.res, o_span, np.float16).astype(np.float32)[idx]
        if d.aux:
            out = out + d.s_aux * _arr(d.aux, o_span, np.float16).astype(np.float32)[idx]
        O = _arr(d.out, o_span, np.float16)
        O[idx] = out.astype(np.float16)
        self.calls.append((d.bx, d.by, d.bz, d.bn))
        return 0

    def wiw_last_error(self):
        return b""


@pytest.fixture()
def fake(monkeypatch):
    lib = FakeLib()
    monkeypatch.setattr(L, "require_device", lambda: lib)
    monkeypatch.setattr(L, "stream_ptr", lambda: None)
    monkeypatch.setattr(ops, "_chk", lambda x: None)
    return lib


def rnd(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def close(y, ref, tol=4e-3):
    err = (y.float() - ref).abs().max().item()
    assert err <= tol * (ref.abs().max().item() + 1e-6), err


def test_linear_bias_residual_aux_table(fake):
    m, k, n = 300, 128, 160
    x = rnd(m, k, seed=1).half()
    w, b = rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3)
    gw = ops.prep_linear(w, b, "cpu")
    res, aux = rnd(m, n, seed=4).half(), rnd(m, n, seed=5).half()
    tab = rnd(5, n, seed=6).half()
    y = ops.linear(x, gw, res=res, aux=aux, tab=tab, tab_div=13, tab_mod=5, s_acc=0.7, s_res=0.3, s_aux=0.25)
    rows = (torch.arange(m) // 13) % 5
    ref = 0.7 * (F.linear(x.float(), w.half().float(), b) + tab.float()[rows]) + 0.3 * res.float() + 0.25 * aux.float()
    close(y, ref)


def test_geglu_interleave(fake):
    m, c = 130, 64
    x = rnd(m, c, seed=1).half()
    w, b = rnd(8 * c, c, seed=2, scale=c ** -0.5), rnd(8 * c, seed=3)
    y = ops.linear(x, ops.prep_linear(w, b, "cpu", geglu=True))
    val, gate = F.linear(x.float(), w.half().float(), b).chunk(2, dim=-1)
    close(y, val * F.gelu(gate))


@pytest.mark.parametrize("n,h,w,cin,cout", [(3, 9, 16, 64, 64), (2, 5, 8, 64, 96), (5, 18, 32, 64, 32), (1, 2, 4, 128, 64)])
def test_conv3x3_taps_and_multi_image_tiles(fake, n, h, w, cin, cout):
    x = rnd(n, h, w, cin, seed=1).half()
    wt, b = rnd(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5), rnd(cout, seed=3)
    gw = ops.prep_conv(wt, b, "cpu")
    tab = rnd(n, gw.n, seed=4).half()
    res = rnd(n, h, w, gw.n, seed=5).half()
    y = ops.conv3x3(x, gw, res=res, tab=tab, tab_div=h * w, tab_mod=n)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float(), b, padding=1).permute(0, 2, 3, 1)
    ref = ref + tab.float()[:, None, None, :cout] + res.float()[..., :cout]
    close(y[..., :cout], ref)
    bx, by, bz, _ = fake.calls[-1]
    assert (bx, by, bz) == ops._box3(w, h, n)


def test_conv_in_small_cin_padded(fake):
    n, h, w, cin, cout = 2, 8, 16, 8, 64
    x = torch.zeros(n, h, w, 64, dtype=torch.float16)
    x[..., :cin] = rnd(n, h, w, cin, seed=1).half()
    wt, b = rnd(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5), rnd(cout, seed=3)
    y = ops.conv3x3(x, ops.prep_conv(wt, b, "cpu"))
    ref = F.conv2d(x[..., :cin].float().permute(0, 3, 1, 2), wt.half().float(), b, padding=1).permute(0, 2, 3, 1)
    close(y[..., :cout], ref)


@pytest.mark.parametrize("n,h,w,c,cout", [(2, 8, 16, 64, 64), (3, 18, 32, 64, 64), (1, 2, 2, 64, 64)])
def test_conv3x3_stride2_space_to_depth(fake, n, h, w, c, cout):
    x = rnd(n, h, w, c, seed=1).half()
    wt, b = rnd(cout, c, 3, 3, seed=2, scale=(9 * c) ** -0.5), rnd(cout, seed=3)
    y = ops.conv3x3_s2(x, ops.prep_conv(wt, b, "cpu"))
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float(), b, stride=2, padding=1).permute(0, 2, 3, 1)
    close(y[..., :cout], ref)


@pytest.mark.parametrize("b_,f,s,c", [(2, 5, 144, 64), (1, 14, 40, 64), (2, 8, 2, 64)])
def test_tconv3_frame_taps_and_blend(fake, b_, f, s, c):
    x = rnd(b_, f, s, c, seed=1).half()
    wt, b = rnd(c, c, 3, 1, 1, seed=2, scale=(3 * c) ** -0.5), rnd(c, seed=3)
    res = rnd(b_, f, s, c, seed=4).half()
    tab = rnd(b_ * f, c, seed=5).half()
    y = ops.tconv3(x, ops.prep_conv(wt, b, "cpu"), res=res, tab=tab, tab_div=s, tab_mod=b_ * f, s_acc=0.3, s_res=1.0)
    xr = x.float().permute(0, 3, 1, 2).unsqueeze(-1)
    ref = F.conv3d(xr, wt.half().float(), b, padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1)
    ref = 0.3 * (ref + tab.float().view(b_, f, 1, c)) + res.float()
    close(y, ref)
    assert fake.calls[-1][:2] == ops._tbox(s, f)
