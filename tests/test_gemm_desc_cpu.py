"""Host logic of the implicit-GEMM wrappers, checked WITHOUT a GPU.

`ops.linear / conv3x3 / conv3x3_s2 / tconv3` only build a `wiw_gemm_desc` (views, strides, filter-tap tables, weight repacking,
GEGLU interleave, epilogue operands) and hand it to the C ABI.  Here the C-ABI call is replaced by a numpy emulator that
executes the *documented meaning* of the descriptor (include/wiw.h) on host memory -- test infrastructure, never the
product -- and the result is compared with torch's own conv / linear.  What this pins on the CPU: the tap tables (zero
padding, the space-to-depth view of the stride-2 conv, the frame taps of the temporal conv), weight layouts, bias / table /
residual / aux / AlphaBlender semantics and the tile-shape choices staying inside the descriptor contract."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wiw_b200  # noqa: F401
from wiw_b200 import lib as L
from wiw_b200 import ops


import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import cabi_emulator  # noqa: E402


@pytest.fixture()
def fake(monkeypatch):
    return cabi_emulator.install(monkeypatch)


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


@pytest.mark.parametrize("n,h,w,c,cout", [(2, 8, 16, 64, 64), (1, 18, 32, 64, 32)])
def test_conv3x3_stride2_asymmetric_padding(fake, n, h, w, c, cout):
    """The VAE encoder's downsampler: F.pad(x, (0, 1, 0, 1)) then a stride-2 conv without padding (downsampling.py)."""
    x = rnd(n, h, w, c, seed=1).half()
    wt, b = rnd(cout, c, 3, 3, seed=2, scale=(9 * c) ** -0.5), rnd(cout, seed=3)
    y = ops.conv3x3_s2(x, ops.prep_conv(wt, b, "cpu"), asym=True)
    xp = F.pad(x.float().permute(0, 3, 1, 2), (0, 1, 0, 1))
    ref = F.conv2d(xp, wt.half().float(), b, stride=2).permute(0, 2, 3, 1)
    close(y[..., :cout], ref)
