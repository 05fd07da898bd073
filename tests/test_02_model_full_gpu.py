"""BASELINE.json configs[0] on the FULL architecture: 256x256 px, 8 frames, 10 Euler steps with CFG, the
1.53 B-parameter UNet (320,640,1280,1280 / heads 5,10,20,20) and the full temporal-VAE decoder, on the
sm_100a kernels (through the C ABI), against the golden output of the reference's own vendored
StableVideoDiffusionPipeline run on the CPU in fp32 (tests/golden/rollout_cfg1_full.npz, written by
oracle/make_golden_full.py in the build container; weights are regenerated here from the same PCG64 seeds).

Tolerance: fp16 storage / fp32 accumulation against an fp32 reference over 10 steps --
final latents and frames PSNR >= 45 dB (PSNR as evaluation/FVD/calculate_psnr.py:6-15), the north-star bar."""
import math
import os

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import ops
from wiw_b200 import weights as W
from wiw_b200.config import FULL_VAE, UNetConfig
from wiw_b200.engine import UNetEngine, VAEDecoderEngine
from wiw_b200.pipeline import Rollout
from wiw_b200.scheduler import EulerSchedule
from oracle import svd_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def psnr_range(a: torch.Tensor, b: torch.Tensor) -> float:
    rng = float(b.max() - b.min())
    mse = float((a - b).pow(2).mean())
    return 100.0 if mse < 1e-20 else 20 * math.log10(rng / math.sqrt(mse))


def test_cfg1_full_architecture_rollout_vs_reference_pipeline(golden_dir):
    g = np.load(os.path.join(golden_dir, "rollout_cfg1_full.npz"))
    n, F = int(g["num_steps"]), g["latents_raw"].shape[1]
    ucfg = UNetConfig(num_frames=F, action_input_channel=F)
    unet = UNetEngine(ucfg, W.random_init(W.unet_param_spec(ucfg), int(g["seed_unet"])), DEV)
    vae = VAEDecoderEngine(FULL_VAE, W.random_init(W.vae_decoder_param_spec(FULL_VAE), int(g["seed_vae"])), DEV)
    ro = Rollout(unet, vae, DEV)
    sched = EulerSchedule(n)
    lat0 = (torch.from_numpy(g["latents_raw"]).half() * sched.init_noise_sigma)[0]
    il = torch.from_numpy(g["image_latents"])[1]
    ie = torch.from_numpy(g["image_embeddings"])[1]
    assert float(torch.from_numpy(g["image_latents"])[0].abs().max()) == 0.0            # CFG row is zeros (pipeline :248)
    lat = ro.denoise(lat0, il, ie, torch.from_numpy(g["added_action_ids"])[0], num_steps=n)
    u8, f32 = ro.decode(lat, want_f32=True)
    torch.cuda.synchronize()
    final = ops.nhwc_to_nchw(lat, 4).float().cpu()
    ref_lat = torch.from_numpy(g["final_latents"])[0]
    assert torch.isfinite(final).all()
    p_lat = psnr_range(final, ref_lat)
    ref_frames = g["frames_f16"][0].astype(np.float32)
    frames = (f32.cpu() / 2 + 0.5).clamp(0, 1).numpy()
    p = O.psnr(frames, ref_frames)
    p8 = O.psnr(u8.cpu().numpy().astype(np.float32) / 255.0, ref_frames)
    from conftest import record_parity
    record_parity("cfg1_full_arch_rollout", latents_psnr_db=p_lat, frames_psnr_db=p, frames_u8_psnr_db=p8)
    print(f"cfg#1 full arch: latents {p_lat:.1f} dB, frames {p:.1f} dB, uint8 frames {p8:.1f} dB")
    assert p_lat > 45.0, f"latents psnr {p_lat:.2f}"
    assert p >= 45.0, f"frame PSNR {p:.2f} dB < 45"
    assert p8 >= 45.0, f"uint8 frame PSNR {p8:.2f} dB < 45"
