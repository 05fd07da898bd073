"""Parity at BASELINE configs[1]'s FULL size against the reference's own pipeline (tests/golden/rollout_northstar.npz, written by
oracle/make_golden_northstar_rollout.py from the vendored StableVideoDiffusionPipeline.__call__ in fp32 on the CPU):
576x1024 px (72x128 latents), 14 frames, CFG, the 1.53 B-parameter UNet and the full temporal VAE, 3 Euler steps.

  * test_full_size_vae_decode  (SURVEY 8a a17): the reference's final latents -> sm_100a temporal-VAE decoder -> frames, compared
    with the reference's decoded frames (two strided sub-samples of different phase + per-frame means of the full frames).
  * test_full_size_rollout     (SURVEY 8a a4): the whole denoise loop + decode from the injected initial latents.

Tolerance: fp16 storage / fp32 accumulation against an fp32 reference -- PSNR >= 45 dB on latents and frames (PSNR as
evaluation/FVD/calculate_psnr.py:6-15; the north-star bar); margins are recorded in gpurun_out/parity.jsonl."""
import math
import os

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import ops
from wiw_b200 import weights as W
from wiw_b200.config import FULL_UNET, FULL_VAE
from wiw_b200.engine import UNetEngine, VAEDecoderEngine
from wiw_b200.pipeline import Rollout
from wiw_b200.scheduler import EulerSchedule

pytestmark = pytest.mark.gpu
DEV = "cuda"


def psnr01(a: np.ndarray, b: np.ndarray) -> float:
    """evaluation/FVD/calculate_psnr.py:6-15 on [0, 1] images."""
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 100.0 if mse < 1e-10 else 20 * math.log10(1.0 / math.sqrt(mse))


def psnr_range(a: torch.Tensor, b: torch.Tensor) -> float:
    rng = float(b.max() - b.min())
    mse = float((a - b).pow(2).mean())
    return 100.0 if mse < 1e-20 else 20 * math.log10(rng / math.sqrt(mse))


def _golden(golden_dir):
    path = os.path.join(golden_dir, "rollout_northstar.npz")
    if not os.path.exists(path):
        pytest.skip("north-star rollout golden not generated")
    return np.load(path)


def _compare_frames(g, f32: torch.Tensor, u8: torch.Tensor, tag: str):
    frames = (f32 / 2 + 0.5).clamp(0, 1)                                             # postprocess_video (pipeline :630-632)
    s4 = frames[:, :, ::4, ::4].cpu().numpy()
    s8 = frames[:, :, 1::8, 3::8].cpu().numpy()
    means = frames.mean(dim=(-1, -2)).cpu().numpy()
    p4, p8 = psnr01(s4, g["frames_s4"].astype(np.float32)), psnr01(s8, g["frames_s8_p13"].astype(np.float32))
    pu = psnr01(u8[:, :, ::4, ::4].cpu().numpy().astype(np.float32) / 255.0, g["frames_s4"].astype(np.float32))
    dmean = float(np.abs(means - g["frame_means"]).max())
    from conftest import record_parity
    record_parity(tag, frames_s4_psnr_db=p4, frames_s8_psnr_db=p8, frames_u8_psnr_db=pu, frame_mean_abs_err=dmean)
    print(f"{tag}: frames {p4:.1f} / {p8:.1f} dB, uint8 {pu:.1f} dB, max |frame-mean error| {dmean:.2e}")
    assert p4 >= 45.0 and p8 >= 45.0, f"frame PSNR {p4:.2f} / {p8:.2f} dB < 45"
    assert pu >= 45.0, f"uint8 frame PSNR {pu:.2f} dB < 45"
    assert dmean < 2e-3


def test_full_size_vae_decode(golden_dir):
    g = _golden(golden_dir)
    vae = VAEDecoderEngine(FULL_VAE, W.random_init(W.vae_decoder_param_spec(FULL_VAE), int(g["seed_vae"])), DEV)
    ro = Rollout.__new__(Rollout)
    ro.vae, ro.device = vae, torch.device(DEV)
    lat = torch.from_numpy(g["final_latents"])[0].half().to(DEV)                     # [F, 4, h, w], the decoder's exact input
    u8, f32 = Rollout.decode(ro, ops.nchw_to_nhwc(lat.contiguous(), 4), want_f32=True)
    torch.cuda.synchronize()
    assert u8.shape == (14, 3, 576, 1024) and torch.isfinite(f32).all()
    _compare_frames(g, f32, u8, "northstar_vae_decode_full_size")


def test_full_size_rollout(golden_dir):
    g = _golden(golden_dir)
    n, F, h, w = int(g["num_steps"]), 14, 72, 128
    unet = UNetEngine(FULL_UNET, W.random_init(W.unet_param_spec(FULL_UNET), int(g["seed_unet"])), DEV)
    vae = VAEDecoderEngine(FULL_VAE, W.random_init(W.vae_decoder_param_spec(FULL_VAE), int(g["seed_vae"])), DEV)
    ro = Rollout(unet, vae, DEV)
    sched = EulerSchedule(n)
    raw = torch.from_numpy(np.random.Generator(np.random.PCG64(int(g["seed_latents"]))).standard_normal((1, F, 4, h, w), dtype=np.float32))
    lat0 = (raw.half() * sched.init_noise_sigma)[0]
    il = torch.from_numpy(g["image_latents"])
    assert float(il[0].abs().max()) == 0.0                                           # CFG row is zeros (pipeline :248)
    lat = ro.denoise(lat0, il[1], torch.from_numpy(g["image_embeddings"])[1], torch.from_numpy(g["added_action_ids"])[0], num_steps=n)
    u8, f32 = ro.decode(lat, want_f32=True)
    torch.cuda.synchronize()
    final = ops.nhwc_to_nchw(lat, 4).float().cpu()
    assert torch.isfinite(final).all()
    p_lat = psnr_range(final, torch.from_numpy(g["final_latents"])[0])
    from conftest import record_parity
    record_parity("northstar_rollout_3_steps", latents_psnr_db=p_lat)
    print(f"north-star 3-step rollout: final latents {p_lat:.1f} dB")
    assert p_lat > 45.0, f"latents psnr {p_lat:.2f}"
    _compare_frames(g, f32, u8, "northstar_rollout_3_steps_frames")
