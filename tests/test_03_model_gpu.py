"""Model-level GPU parity: the sm_100a UNet / temporal-VAE / rollout engines (called through the
C ABI) against the CPU oracle on the same seeded inputs, and against the golden outputs of the
reference's vendored fork (tests/golden/*.npz).

Tolerances (fp16 storage + fp32 accumulation vs an fp32 oracle; the reference itself runs fp16):
  single UNet forward / VAE decode: relative RMS error <= 1e-2 and PSNR >= 45 dB on the output range
  full rollout frames: PSNR >= 45 dB vs the reference frames (the north-star bar)."""
import math
import os

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import ops
from wiw_b200 import weights as W
from wiw_b200.config import TINY_UNET, TINY_VAE
from wiw_b200.engine import UNetEngine, VAEDecoderEngine, sinusoid
from wiw_b200.pipeline import Rollout
from wiw_b200.scheduler import EulerSchedule
from oracle import svd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
SEED_UNET, SEED_VAE = 11, 12
UCFG = O.UNetCfg(block_out_channels=TINY_UNET.block_out_channels, num_attention_heads=TINY_UNET.num_attention_heads,
                 cross_attention_dim=TINY_UNET.cross_attention_dim, num_frames=TINY_UNET.num_frames,
                 action_input_channel=TINY_UNET.action_input_channel)
VCFG = O.VAECfg(block_out_channels=TINY_VAE.block_out_channels)


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    return float(((a - b).pow(2).mean() / b.pow(2).mean()).sqrt())


def psnr_range(a: torch.Tensor, b: torch.Tensor) -> float:
    rng = float(b.max() - b.min())
    mse = float((a - b).pow(2).mean())
    return 100.0 if mse < 1e-20 else 20 * math.log10(rng / math.sqrt(mse))


@pytest.fixture(scope="module")
def unet_sd():
    return W.random_init(W.unet_param_spec(TINY_UNET), SEED_UNET)


@pytest.fixture(scope="module")
def vae_sd():
    return W.random_init(W.vae_decoder_param_spec(TINY_VAE), SEED_VAE)


@pytest.fixture(scope="module")
def unet(unet_sd):
    return UNetEngine(TINY_UNET, unet_sd, DEV)


@pytest.fixture(scope="module")
def vae(vae_sd):
    return VAEDecoderEngine(TINY_VAE, vae_sd, DEV)


def _unet_inputs(golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_tiny.npz"))
    return g, torch.from_numpy(g["sample"]), torch.from_numpy(g["enc"]), torch.from_numpy(g["added_action_ids"])


def test_unet_forward_vs_golden_and_oracle(golden_dir, unet, unet_sd):
    g, sample, enc, aids = _unet_inputs(golden_dir)
    nb, F, _, h, w = sample.shape
    unet.set_condition(enc[1], 0.02, aids[0])
    x_in = ops.nchw_to_nhwc(sample.flatten(0, 1).half().to(DEV).contiguous(), 64)           # [nb*F, h, w, 64]
    t_emb = sinusoid(torch.tensor([float(g["timestep"])]), TINY_UNET.block_out_channels[0]).half().to(DEV)
    taps = {}
    v = unet.forward(x_in, t_emb, taps=taps)
    torch.cuda.synchronize()
    out = ops.nhwc_to_nchw(v, 4).float().cpu().view(nb, F, 4, h, w)
    ref = torch.from_numpy(g["out"])
    # where does the error come from: per-block taps against the reference's own activations
    for k in ("down.0.0", "mid", "up.3.2"):
        t = taps[k]
        got = ops.nhwc_to_nchw(t, t.shape[-1]).float().cpu()
        r = torch.from_numpy(g["tap." + k])
        assert rel_rms(got, r) < 1e-2, f"tap {k}: rel rms {rel_rms(got, r):.4g}"
    assert torch.isfinite(out).all()
    from conftest import record_parity
    record_parity("tiny_unet_forward_vs_golden", rel_rms=rel_rms(out, ref), psnr_db=psnr_range(out, ref))
    assert rel_rms(out, ref) < 1e-2, f"unet out rel rms {rel_rms(out, ref):.4g}"
    assert psnr_range(out, ref) > 45.0
    # and against the oracle evaluated on the fp16-rounded inputs (what the kernels actually saw)
    with torch.no_grad():
        orc = O.unet_forward(unet_sd, UCFG, sample.half().float(), torch.tensor(float(g["timestep"])), enc.half().float(),
                             torch.from_numpy(g["added_time_ids"]), aids)
    assert rel_rms(out, orc) < 1e-2


def test_vae_decode_vs_golden(golden_dir, vae):
    g = np.load(os.path.join(golden_dir, "vae_tiny.npz"))
    z = torch.from_numpy(g["z"])
    F, _, h, w = z.shape
    zin = ops.nchw_to_nhwc(z.half().to(DEV).contiguous(), 64)
    f32 = torch.empty(F, 3, 8 * h, 8 * w, dtype=torch.float32, device=DEV)
    u8 = torch.empty(F, 3, 8 * h, 8 * w, dtype=torch.uint8, device=DEV)
    vae.decode(zin, F, out_u8=u8, out_f32=f32)
    torch.cuda.synchronize()
    ref = torch.from_numpy(g["out"])
    got = f32.cpu()
    assert torch.isfinite(got).all()
    assert rel_rms(got, ref) < 1e-2, f"vae rel rms {rel_rms(got, ref):.4g}"
    q = ((ref / 2 + 0.5).clamp(0, 1) * 255).round()
    assert float((u8.cpu().float() - q).abs().max()) <= 1.0          # quantisation may flip at a rounding edge


def test_rollout_vs_reference_pipeline(golden_dir, unet, vae):
    """4 Euler steps with CFG + decode vs the vendored StableVideoDiffusionPipeline: PSNR >= 45 dB."""
    g = np.load(os.path.join(golden_dir, "rollout_tiny.npz"))
    n = int(g["num_steps"])
    sched = EulerSchedule(n)
    lat0 = (torch.from_numpy(g["latents_raw"]).half() * sched.init_noise_sigma)[0]          # fp16 * 0-dim fp32 (pipeline :361)
    il = torch.from_numpy(g["image_latents"])[1]
    ie = torch.from_numpy(g["image_embeddings"])[1]
    ro = Rollout(unet, vae, DEV)
    trace = []
    lat = ro.denoise(lat0, il, ie, torch.from_numpy(g["added_action_ids"])[0], num_steps=n, trace=trace)
    u8, f32 = ro.decode(lat, want_f32=True)
    torch.cuda.synchronize()
    final = ops.nhwc_to_nchw(lat, 4).float().cpu()
    ref_lat = torch.from_numpy(g["final_latents"])[0]
    assert torch.isfinite(final).all()
    assert psnr_range(final, ref_lat) > 45.0, f"latents psnr {psnr_range(final, ref_lat):.2f}"
    frames = (f32.cpu() / 2 + 0.5).clamp(0, 1).numpy()
    p = O.psnr(frames, g["frames"][0])
    assert p >= 45.0, f"frame PSNR {p:.2f} dB < 45"
    p8 = O.psnr(u8.cpu().numpy().astype(np.float32) / 255.0, g["frames"][0])
    from conftest import record_parity
    record_parity("tiny_rollout_vs_reference_pipeline", latents_psnr_db=psnr_range(final, ref_lat), frames_psnr_db=p, frames_u8_psnr_db=p8)
    assert p8 >= 45.0, f"uint8 frame PSNR {p8:.2f} dB < 45"


def test_scheduler_index_arithmetic_bit_exact(golden_dir):
    g = np.load(os.path.join(golden_dir, "scheduler.npz"))
    for n in (10, 25, 30):
        s = EulerSchedule(n)
        assert s.sigmas.numpy().tobytes() == g[f"sigmas_{n}"].tobytes()
        assert s.timesteps.numpy().tobytes() == g[f"timesteps_{n}"].tobytes()
        assert [s.index_for_timestep(t) for t in s.timesteps] == list(range(n))


# ----------------------------------------------------------------------------- conditioning front-end on the sm_100a kernels
def test_vae_encoder_engine_vs_reference_golden(golden_dir):
    """engine.VAEEncoderEngine (SD-VAE encoder + quant_conv on the implicit-GEMM / GroupNorm / attention kernels, fp16 storage,
    fp32 accumulation) against the reference's own fp32 `vae.encode(x).latent_dist.mode()` (tests/golden/frontend_tiny.npz).
    Tolerance: the reference runs this stage in fp32 (force_upcast); fp16 storage between ~25 layers -> relative RMS <= 1e-2 and
    PSNR >= 45 dB on the latent range."""
    from wiw_b200.engine import VAEEncoderEngine
    from wiw_b200.frontend import vae_encoder_param_spec
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    c = TINY_VAE
    esd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), SEED_VAE + 1)
    enc = VAEEncoderEngine(c, esd, DEV)
    got = enc.encode_mode(torch.from_numpy(g["image"])).float().cpu()
    ref = torch.from_numpy(g["latent_mode"])
    torch.cuda.synchronize()
    assert got.shape == ref.shape and torch.isfinite(got).all()
    r, p = rel_rms(got, ref), psnr_range(got, ref)
    from conftest import record_parity
    record_parity("vae_encoder_engine_tiny_vs_reference_fp32", rel_rms=r, psnr_db=p)
    assert r < 1e-2 and p > 45.0, (r, p)


def test_vae_encoder_engine_full_size_vs_fp32_path():
    """Same engine at the deployed size (576x1024 px, full SD-VAE encoder) against the fp32 PyTorch path the worker uses by
    default (frontend.vae_encode_mode = the reference's arithmetic, checked against the golden on the CPU)."""
    from wiw_b200.config import FULL_VAE
    from wiw_b200.engine import VAEEncoderEngine
    from wiw_b200.frontend import vae_encode_mode, vae_encoder_param_spec
    c = FULL_VAE
    esd = W.random_init_on_device(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), 2, DEV)
    img = torch.from_numpy(np.random.Generator(np.random.PCG64(9)).uniform(-1, 1, (1, 3, 576, 1024)).astype(np.float32)).to(DEV)
    torch.backends.cudnn.allow_tf32 = False
    ref = vae_encode_mode({k: v.float() for k, v in esd.items()}, img).float().cpu()
    torch.backends.cudnn.allow_tf32 = True
    got = VAEEncoderEngine(c, esd, DEV).encode_mode(img).float().cpu()
    torch.cuda.synchronize()
    assert got.shape == (1, 4, 72, 128) and torch.isfinite(got).all()
    r, p = rel_rms(got, ref), psnr_range(got, ref)
    from conftest import record_parity
    record_parity("vae_encoder_engine_full_size_vs_fp32_path", rel_rms=r, psnr_db=p)
    assert r < 1e-2 and p > 45.0, (r, p)
