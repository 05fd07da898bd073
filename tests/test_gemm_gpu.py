"""GPU parity tests of the tcgen05 implicit-GEMM kernel (wiw_conv_gemm) against a plain
PyTorch fp32 reference of the same op on the same fp16-rounded inputs.
Tolerance: the kernel accumulates in fp32 and rounds once to fp16, so
|err| <= 2^-10 * |y| (half an fp16 ulp, relative) + 2e-3 * scale slack for accumulation order."""
import pytest
import torch
import torch.nn.functional as F

import wiw_b200  # noqa: F401
from wiw_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def close(y: torch.Tensor, ref: torch.Tensor, what=""):
    y = y.float()
    scale = ref.abs().max().item() + 1e-6
    err = (y - ref).abs()
    tol = 2e-3 * ref.abs() + 2e-3 * scale
    bad = (err > tol).sum().item()
    assert torch.isfinite(y).all(), f"{what}: non-finite output"
    assert bad == 0, f"{what}: {bad} elements out of tolerance, max err {err.max().item():.4g} (scale {scale:.3g})"


@pytest.mark.parametrize("m,k,n", [(300, 128, 160), (128, 64, 32), (4032, 320, 960), (1000, 1280, 320), (17, 64, 256),
                                   (700, 192, 640), (260, 64, 288), (40000, 1280, 640)])   # last three: ragged last N-tile (bn 224 / 160)
def test_linear_bias(m, k, n):
    x = rnd(m, k, seed=1).half().to(DEV)
    w, b = rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3)
    gw = ops.prep_linear(w, b, DEV)
    y = ops.linear(x, gw)
    ref = F.linear(x.float(), w.half().float().to(DEV), b.to(DEV))
    close(y, ref, f"linear {m}x{k}x{n}")


def test_linear_no_bias_unaligned_k_n():
    """k and n not multiples of 64/32 are zero-padded at weight-prep time (cross-attn dim 32, N=4)."""
    m, k, n = 257, 96, 40
    x = torch.zeros(m, 128).half()
    x[:, :k] = rnd(m, k, seed=4).half()
    w = rnd(n, k, seed=5, scale=0.1)
    gw = ops.prep_linear(w, None, DEV)
    assert gw.k == 128 and gw.n == 64
    y = ops.linear(x.to(DEV), gw)
    ref = F.linear(x[:, :k].float().to(DEV), w.half().float().to(DEV))
    close(y[:, :n], ref, "linear padded")
    assert float(y[:, n:].abs().max()) == 0.0


def test_linear_full_epilogue():
    """out = s_acc*(xW^T + b + tab[(row/div)%mod]) + s_res*res + s_aux*aux"""
    m, k, n = 700, 320, 256
    x = rnd(m, k, seed=1).half().to(DEV)
    w, b = rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3)
    res, aux = rnd(m, n, seed=4).half().to(DEV), rnd(m, n, seed=5).half().to(DEV)
    tab = rnd(5, n, seed=6).half().to(DEV)
    gw = ops.prep_linear(w, b, DEV)
    y = ops.linear(x, gw, res=res, aux=aux, tab=tab, tab_div=50, tab_mod=5, s_acc=0.4, s_res=0.4, s_aux=0.6)
    rows = (torch.arange(m, device=DEV) // 50) % 5
    ref = 0.4 * (F.linear(x.float(), w.half().float().to(DEV), b.to(DEV)) + tab.float()[rows]) + 0.4 * res.float() + 0.6 * aux.float()
    close(y, ref, "linear epilogue")


@pytest.mark.parametrize("c", [64, 320])
def test_geglu(c):
    m = 520
    x = rnd(m, c, seed=1).half().to(DEV)
    w, b = rnd(8 * c, c, seed=2, scale=c ** -0.5), rnd(8 * c, seed=3)
    gw = ops.prep_linear(w, b, DEV, geglu=True)
    y = ops.linear(x, gw)
    h = F.linear(x.float(), w.half().float().to(DEV), b.to(DEV))
    val, gate = h.chunk(2, dim=-1)
    close(y, val * F.gelu(gate), f"geglu C={c}")


@pytest.mark.parametrize("n,h,w,cin,cout", [(3, 9, 16, 64, 128), (2, 16, 128, 128, 320), (2, 5, 8, 64, 64), (1, 2, 4, 128, 64), (2, 72, 128, 64, 32), (2, 18, 32, 128, 640),
                                             # M-tiles spanning several images (bz > 1): exact and ragged image counts
                                             (28, 9, 16, 64, 64), (5, 9, 16, 64, 96), (7, 18, 32, 64, 64), (3, 1, 2, 64, 64)])
def test_conv3x3(n, h, w, cin, cout):
    x = rnd(n, h, w, cin, seed=1).half().to(DEV)
    wt, b = rnd(cout, cin, 3, 3, seed=2, scale=(9 * cin) ** -0.5), rnd(cout, seed=3)
    gw = ops.prep_conv(wt, b, DEV)
    tab = rnd(n, gw.n, seed=4).half().to(DEV)
    res = rnd(n, h, w, gw.n, seed=5).half().to(DEV)
    y = ops.conv3x3(x, gw, res=res, tab=tab, tab_div=h * w, tab_mod=n)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float().to(DEV), b.to(DEV), padding=1).permute(0, 2, 3, 1)
    ref = ref + tab.float()[:, None, None, :cout] + res.float()[..., :cout]
    close(y[..., :cout], ref, f"conv3x3 {n}x{h}x{w}x{cin}->{cout}")


def test_conv3x3_small_cin_padded():
    """conv_in: 8 input channels zero-padded to 64."""
    n, h, w, cin, cout = 2, 8, 16, 8, 64
    xs = rnd(n, h, w, cin, seed=1).half()
    x = torch.zeros(n, h, w, 64).half()
    x[..., :cin] = xs
    wt, b = rnd(cout, cin, 3, 3, seed=2, scale=0.1), rnd(cout, seed=3)
    gw = ops.prep_conv(wt, b, DEV)
    y = ops.conv3x3(x.to(DEV), gw)
    ref = F.conv2d(xs.float().to(DEV).permute(0, 3, 1, 2), wt.half().float().to(DEV), b.to(DEV), padding=1).permute(0, 2, 3, 1)
    close(y, ref, "conv_in")


@pytest.mark.parametrize("n,h,w,c,cout", [(2, 8, 16, 64, 64), (3, 36, 64, 128, 128), (1, 2, 2, 64, 64)])
def test_conv3x3_stride2(n, h, w, c, cout):
    x = rnd(n, h, w, c, seed=1).half().to(DEV)
    wt, b = rnd(cout, c, 3, 3, seed=2, scale=(9 * c) ** -0.5), rnd(cout, seed=3)
    gw = ops.prep_conv(wt, b, DEV)
    y = ops.conv3x3_s2(x, gw)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float().to(DEV), b.to(DEV), stride=2, padding=1).permute(0, 2, 3, 1)
    close(y, ref, "conv s2")


@pytest.mark.parametrize("b_,f,s,c", [(2, 5, 144, 64), (1, 14, 128, 128), (2, 8, 2, 64), (2, 14, 144, 64), (1, 14, 576, 64),
                                        (2, 3, 40, 64)])
def test_tconv3(b_, f, s, c):
    x = rnd(b_, f, s, c, seed=1).half().to(DEV)
    wt, b = rnd(c, c, 3, 1, 1, seed=2, scale=(3 * c) ** -0.5), rnd(c, seed=3)
    gw = ops.prep_conv(wt, b, DEV)
    res = rnd(b_, f, s, c, seed=4).half().to(DEV)
    tab = rnd(b_ * f, c, seed=5).half().to(DEV)
    y = ops.tconv3(x, gw, res=res, tab=tab, tab_div=s, tab_mod=b_ * f, s_acc=0.3, s_res=1.0)
    xr = x.float().permute(0, 3, 1, 2).unsqueeze(-1)                     # [B, C, F, S, 1]
    ref = F.conv3d(xr, wt.half().float().to(DEV), b.to(DEV), padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1)
    ref = 0.3 * (ref + tab.float().view(b_, f, 1, c)) + res.float()
    close(y, ref, "tconv3")


def test_conv1x1_matches_linear():
    n, h, w, cin, cout = 2, 4, 8, 128, 64
    x = rnd(n, h, w, cin, seed=1).half().to(DEV)
    wt, b = rnd(cout, cin, 1, 1, seed=2, scale=cin ** -0.5), rnd(cout, seed=3)
    gw = ops.prep_conv(wt, b, DEV)
    y = ops.conv1x1(x, gw)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.half().float().to(DEV), b.to(DEV)).permute(0, 2, 3, 1)
    close(y, ref, "conv1x1")


def test_many_tiles_persistent_loop():
    """More tiles than SMs: each CTA loops over several tiles (both TMEM stages, ring wrap-around)."""
    m, k, n = 128 * 40, 192, 1280
    x = rnd(m, k, seed=1).half().to(DEV)
    w, b = rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3)
    res = rnd(m, n, seed=4).half().to(DEV)
    gw = ops.prep_linear(w, b, DEV)
    y = ops.linear(x, gw, res=res)
    ref = F.linear(x.float(), w.half().float().to(DEV), b.to(DEV)) + res.float()
    close(y, ref, "persistent")


@pytest.mark.parametrize("kind", ["linear_res", "linear_res_aux", "conv_res", "geglu"])
def test_epilogue_is_bit_reproducible(kind):
    """The kernel has no atomics and a fixed summation order, so a launch repeated on the same operands must be bit-identical.
    This is the race detector for the epilogue's shared-memory rings: without the proxy fence between the generic-proxy reads of
    a residual slot and its next TMA overwrite, 1-3 % of launches of these layer classes (odd chunk counts per tile, full-size M)
    picked up a few 16-byte pieces of another chunk's residual (scripts/stress_gemm_res3.py)."""
    reps = 150
    if kind in ("linear_res", "linear_res_aux"):
        m, k, n = 258048, 1280, 320                                  # bn = 160: 5 chunks per tile
        x = rnd(m, k, seed=1).half().to(DEV)
        gw = ops.prep_linear(rnd(n, k, seed=2, scale=k ** -0.5), rnd(n, seed=3), DEV)
        res = rnd(m, n, seed=4).half().to(DEV)
        aux = rnd(m, n, seed=5).half().to(DEV) if kind == "linear_res_aux" else None
        out = torch.empty(m, n, device=DEV, dtype=torch.half)
        fn = (lambda: ops.linear(x, gw, res=res, out=out)) if aux is None else \
             (lambda: ops.linear(x, gw, res=res, aux=aux, s_acc=0.7, s_res=0.7, s_aux=0.3, out=out))
    elif kind == "conv_res":
        x = rnd(28, 72, 128, 320, seed=1).half().to(DEV)
        gw = ops.prep_conv(rnd(320, 320, 3, 3, seed=2, scale=0.02), rnd(320, seed=3), DEV)
        res = rnd(28, 72, 128, 320, seed=4).half().to(DEV)
        out = torch.empty(28, 72, 128, 320, device=DEV, dtype=torch.half)
        fn = lambda: ops.conv3x3(x, gw, res=res, out=out)
    else:
        m, c = 64512, 640
        x = rnd(m, c, seed=1).half().to(DEV)
        gw = ops.prep_linear(rnd(8 * c, c, seed=2, scale=c ** -0.5), rnd(8 * c, seed=3), DEV, geglu=True)
        out = torch.empty(m, 4 * c, device=DEV, dtype=torch.half)
        fn = lambda: ops.linear(x, gw, out=out)
    ref = fn().clone()
    bad = sum(0 if torch.equal(fn(), ref) else 1 for _ in range(reps))
    assert bad == 0, f"{kind}: {bad}/{reps} repeated launches differ from the first"
