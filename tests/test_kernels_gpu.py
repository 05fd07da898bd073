"""GPU parity tests of the non-GEMM kernels (norms, attention, sampler glue, layout helpers)
against plain PyTorch fp32 references on the same fp16-rounded inputs."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import wiw_b200  # noqa: F401
from wiw_b200 import lib as L
from wiw_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


def close(y, ref, what="", rel=2e-3, abs_scale=2e-3):
    y = y.float()
    scale = ref.abs().max().item() + 1e-6
    err = (y - ref).abs()
    assert torch.isfinite(y).all(), f"{what}: non-finite"
    bad = (err > rel * ref.abs() + abs_scale * scale).sum().item()
    assert bad == 0, f"{what}: {bad} bad, max err {err.max().item():.4g}, scale {scale:.3g}"


@pytest.mark.parametrize("samples,rows,c,silu", [(3, 144, 64, True), (28, 576, 320, True), (2, 14 * 144, 960, False),
                                                 (2, 37, 1920, True), (1, 5000, 128, False),
                                                 # many tiles per CTA: the 4-stage TMA ring of the streaming kernels wraps several times
                                                 (2, 60000, 320, True), (3, 20011, 640, False), (1, 9000, 1280, True),
                                                 (2, 300, 2560, True)])
def test_groupnorm(samples, rows, c, silu):
    x = (rnd(samples, rows, c, seed=1) * 2 + 0.7).half().to(DEV)
    g, b = 1 + 0.1 * rnd(c, seed=2), 0.1 * rnd(c, seed=3)
    nw = ops.prep_norm(g, b, DEV)
    y = ops.group_norm(x, nw, samples, 1e-5, silu)
    ref = F.group_norm(x.float().permute(0, 2, 1), 32, g.to(DEV), b.to(DEV), 1e-5)
    if silu:
        ref = F.silu(ref)
    close(y, ref.permute(0, 2, 1), f"gn {samples}x{rows}x{c}")


@pytest.mark.parametrize("m,c", [(1000, 64), (333, 320), (50, 640), (7, 1280), (4100, 128),
                                 (120001, 320), (40000, 640), (20003, 1280), (3000, 2048)])     # ring wraps several times
def test_layernorm(m, c):
    x = (rnd(m, c, seed=1) * 3 + 1).half().to(DEV)
    g, b = 1 + 0.1 * rnd(c, seed=2), 0.1 * rnd(c, seed=3)
    nw = ops.prep_norm(g, b, DEV)
    y = ops.layer_norm(x, nw)
    close(y, F.layer_norm(x.float(), (c,), g.to(DEV), b.to(DEV), 1e-5), f"ln {m}x{c}")
    tab = rnd(5, c, seed=4).half().to(DEV)
    y2 = ops.layer_norm(x, nw, tab=tab, tab_div=13, tab_mod=5)
    rows = (torch.arange(m, device=DEV) // 13) % 5
    xin = (x + tab[rows]).float()           # fp16 add, as the reference does
    close(y2, F.layer_norm(xin, (c,), g.to(DEV), b.to(DEV), 1e-5), f"ln+tab {m}x{c}")


@pytest.mark.parametrize("images,s,c", [(2, 128, 64), (3, 576, 128), (2, 144, 64), (1, 2304, 320), (2, 200, 64), (1, 9216, 64)])
def test_attention_spatial(images, s, c):
    qkv = rnd(images * s, 3 * c, seed=1).half().to(DEV)
    y = ops.attention_spatial(qkv, images, s, c)
    heads = c // 64
    q, k, v = [t.float().view(images, s, heads, 64).transpose(1, 2) for t in qkv.view(images, s, 3 * c).chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(images * s, c)
    close(y, ref, f"attn {images}x{s}x{c}", rel=4e-3, abs_scale=3e-3)


def test_attention_spatial_peaky_rows():
    """Large logits (one dominant key per row, growing along the sequence) exercise the lazy rescale."""
    images, s, c = 1, 1024, 64
    g = torch.Generator().manual_seed(5)
    q = torch.randn(s, 64, generator=g) * 3
    k = torch.randn(s, 64, generator=g) * 3 * torch.linspace(0.2, 2.0, s)[:, None]
    v = torch.randn(s, 64, generator=g)
    qkv = torch.cat([q, k, v], dim=1).half().to(DEV)
    y = ops.attention_spatial(qkv, images, s, c)
    qq, kk, vv = [t.float()[None, None] for t in qkv.chunk(3, dim=-1)]
    ref = F.scaled_dot_product_attention(qq, kk, vv)[0, 0]
    close(y, ref, "attn peaky", rel=6e-3, abs_scale=4e-3)


@pytest.mark.parametrize("b,f,s,c", [(2, 8, 128, 64), (2, 14, 144, 128), (1, 14, 50, 320), (2, 8, 2, 128), (1, 25, 40, 64)])
def test_attention_temporal(b, f, s, c):
    qkv = rnd(b * f * s, 3 * c, seed=1).half().to(DEV)
    y = ops.attention_temporal(qkv, b, f, s, c)
    heads = c // 64
    t = qkv.float().view(b, f, s, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)          # [3, b, s, heads, f, 64]
    ref = F.scaled_dot_product_attention(t[0], t[1], t[2])                          # [b, s, heads, f, 64]
    ref = ref.permute(0, 3, 1, 2, 4).reshape(b * f * s, c)
    close(y, ref, f"tattn {b}x{f}x{s}x{c}")


def test_concat_upsample_transpose_layout():
    a, b = rnd(37, 64, seed=1).half().to(DEV), rnd(37, 128, seed=2).half().to(DEV)
    assert torch.equal(ops.concat_channels(a, b), torch.cat([a, b], dim=1))
    x = rnd(2, 3, 5, 64, seed=3).half().to(DEV)
    ref = F.interpolate(x.permute(0, 3, 1, 2).float(), scale_factor=2.0, mode="nearest").permute(0, 2, 3, 1).half()
    assert torch.equal(ops.upsample_nearest2x(x), ref)
    t = rnd(3, 70, 40, seed=4).half().to(DEV)
    assert torch.equal(ops.transpose(t), t.transpose(1, 2).contiguous())
    n = rnd(2, 5, 6, 7, seed=5).half().to(DEV)                 # NCHW
    cl = ops.nchw_to_nhwc(n, 8)
    assert torch.equal(cl[..., :5], n.permute(0, 2, 3, 1)) and float(cl[..., 5:].abs().max()) == 0
    assert torch.equal(ops.nhwc_to_nchw(cl, 5), n)


def test_softmax_rows():
    s = (rnd(50, 1024, seed=1) * 4).half().to(DEV)
    ref = torch.softmax(s.float() * 0.3, dim=-1)
    close(ops.softmax_rows_(s.clone(), 0.3), ref, "softmax", rel=3e-3, abs_scale=1e-3)


def test_sampler_glue_bit_exact():
    """wiw_sampler_prepare / wiw_sampler_step vs the reference's tensor ops evaluated by torch on the
    CPU in the same dtypes (scheduler :313-314, :635-673; pipeline :589-608): bit-exact."""
    lib = L.require_device()
    F_, hw = 5, 96
    lat = (rnd(F_, hw, 4, seed=1) * 300).half()
    img = rnd(hw, 4, seed=2).half()
    sigma, sigma_next = torch.tensor(421.569122), torch.tensor(322.453674)
    div = (sigma ** 2 + 1) ** 0.5
    xin = torch.zeros(2, F_, hw, 64, dtype=torch.float16, device=DEV)
    lat_d, img_d = lat.to(DEV), img.to(DEV)
    L.check(lib.wiw_sampler_prepare(lat_d.data_ptr(), img_d.data_ptr(), xin.data_ptr(), F_, hw, float(div), L.stream_ptr()))
    scaled = lat / div                                       # fp16 tensor / 0-dim fp32 tensor -> fp16
    exp = torch.zeros(2, F_, hw, 64, dtype=torch.float16)
    exp[:, :, :, :4] = scaled
    exp[1, :, :, 4:8] = img
    assert torch.equal(xin.cpu(), exp)

    v = rnd(2, F_, hw, 32, seed=3).half()
    gs = torch.linspace(1.0, 3.0, F_).half()
    c_out = -sigma / (sigma ** 2 + 1) ** 0.5
    sig2p1 = sigma ** 2 + 1
    dt = sigma_next - sigma
    v_d, gs_d = v.to(DEV), gs.to(DEV)
    L.check(lib.wiw_sampler_step(v_d.data_ptr(), lat_d.data_ptr(), gs_d.data_ptr(), F_, hw, 32, float(c_out), float(sig2p1),
                                 float(sigma), float(dt), L.stream_ptr()))
    vu, vc = v[0, ..., :4], v[1, ..., :4]
    noise_pred = vu + gs[:, None, None] * (vc - vu)          # fp16 tensor ops, rounded at each step
    sample = lat.float()
    pred = noise_pred * c_out + (sample / sig2p1)            # fp16 * 0-dim fp32 -> fp16, then + fp32
    deriv = (sample - pred) / sigma
    prev = (sample + deriv * dt).half()
    assert torch.equal(lat_d.cpu(), prev)


def test_vae_finish_and_u8():
    lib = L.require_device()
    F_, hw, cs = 4, 200, 32
    x = rnd(F_, hw, cs, seed=1).half().to(DEV)
    w, b = rnd(3, 3, 3, seed=2) * 0.3, rnd(3, seed=3) * 0.1
    w_d, b_d = w.contiguous().to(DEV), b.to(DEV)
    u8 = torch.empty(F_, 3, hw, dtype=torch.uint8, device=DEV)
    f32 = torch.empty(F_, 3, hw, dtype=torch.float32, device=DEV)
    L.check(lib.wiw_vae_finish(x.data_ptr(), w_d.data_ptr(), b_d.data_ptr(), u8.data_ptr(), f32.data_ptr(), F_, hw, cs, L.stream_ptr()))
    xin = x[..., :3].float().permute(2, 0, 1)[None, :, :, :, None]          # [1, 3, F, hw, 1]
    ref = F.conv3d(xin, w.to(DEV)[:, :, :, None, None], b_d, padding=(1, 0, 0))[0, :, :, :, 0].permute(1, 0, 2)
    close(f32, ref, "time_conv_out")
    q = ((f32 / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
    assert torch.equal(u8, q)


@pytest.mark.parametrize("H,W,oh,ow", [(576, 1024, 480, 480), (64, 128, 40, 48), (30, 40, 50, 64)])
def test_gpu_resize_is_pil_exact(H, W, oh, ow):
    """wiw_resize_bicubic_u8 == PIL Image.resize(BICUBIC) byte for byte, then the reference's requantisation chain."""
    from PIL import Image
    from wiw_b200.resize import GpuResizer
    rng = np.random.default_rng(H + W)
    planes = rng.integers(0, 256, size=(2, 3, H, W), dtype=np.uint8)
    out = GpuResizer(H, W, oh, ow, DEV, requant=False)(torch.from_numpy(planes).to(DEV)).cpu().numpy()
    for i in range(2):
        ref = np.asarray(Image.fromarray(planes[i].transpose(1, 2, 0)).resize((ow, oh), Image.BICUBIC)).transpose(2, 0, 1)
        assert np.array_equal(out[i], ref)
    out2 = GpuResizer(H, W, oh, ow, DEV, requant=True)(torch.from_numpy(planes).to(DEV)).cpu().numpy()
    f = out.astype(np.float32) / np.float32(255)
    assert np.array_equal(out2, (np.clip(f, 0, 1) * 255).astype(np.uint8))


@pytest.mark.parametrize("H,W,oh,ow", [(480, 480, 576, 1024), (384, 384, 576, 1024), (40, 48, 64, 128)])
def test_gpu_resize_lanczos_is_pil_exact(H, W, oh, ow):
    """The request side: the same integer resampler with Pillow's LANCZOS window == PIL Image.resize(LANCZOS) byte for byte
    (video_processor.preprocess of a solver-sized observation, DP/image_processor.py:501)."""
    from PIL import Image
    from wiw_b200.resize import GpuResizer
    planes = np.random.default_rng(H * 3 + W).integers(0, 256, size=(2, 3, H, W), dtype=np.uint8)
    out = GpuResizer(H, W, oh, ow, DEV, requant=False, filt="lanczos")(torch.from_numpy(planes).to(DEV)).cpu().numpy()
    for i in range(2):
        ref = np.asarray(Image.fromarray(planes[i].transpose(1, 2, 0)).resize((ow, oh), Image.LANCZOS)).transpose(2, 0, 1)
        assert np.array_equal(out[i], ref)
