"""Pins the CPU oracle (oracle/svd_oracle.py) to outputs of the reference's vendored
diffusers fork frozen in tests/golden/ by oracle/make_golden.py.  CPU only."""
import os

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import weights as W
from wiw_b200.config import TINY_UNET, TINY_VAE
from oracle import svd_oracle as O

SEED_UNET, SEED_VAE = 11, 12
UCFG = O.UNetCfg(block_out_channels=TINY_UNET.block_out_channels, num_attention_heads=TINY_UNET.num_attention_heads,
                 cross_attention_dim=TINY_UNET.cross_attention_dim, num_frames=TINY_UNET.num_frames,
                 action_input_channel=TINY_UNET.action_input_channel)
VCFG = O.VAECfg(block_out_channels=TINY_VAE.block_out_channels)


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.fixture(scope="module")
def unet_sd():
    return W.random_init(W.unet_param_spec(TINY_UNET), SEED_UNET)


@pytest.fixture(scope="module")
def vae_sd():
    return W.random_init(W.vae_decoder_param_spec(TINY_VAE), SEED_VAE)


@pytest.mark.parametrize("n", [10, 25, 30])
def test_scheduler_tables_bit_exact(golden_dir, n):
    g = _load(golden_dir, "scheduler.npz")
    s = O.EulerSchedule(n)
    assert s.sigmas.numpy().tobytes() == g[f"sigmas_{n}"].tobytes()
    assert s.timesteps.numpy().tobytes() == g[f"timesteps_{n}"].tobytes()
    assert np.float32(s.init_noise_sigma.item()).tobytes() == g[f"init_noise_sigma_{n}"].tobytes()
    idx = np.array([s.index_for_timestep(t) for t in s.timesteps])
    assert (idx == g[f"index_{n}"]).all() and (idx == np.arange(n)).all()


def test_scheduler_known_answers():
    """Known answers recorded in BASELINE.md section 5."""
    s = O.EulerSchedule(25)
    assert abs(s.init_noise_sigma.item() - 700.000732421875) < 1e-9
    np.testing.assert_allclose(s.sigmas[:4].numpy(), [700.0, 545.729248, 421.569122, 322.453674], rtol=1e-6)
    np.testing.assert_allclose(s.timesteps[:3].numpy(), [1.637770, 1.575531, 1.510996], rtol=1e-6)
    assert float(s.sigmas[-1]) == 0.0 and abs(float(s.sigmas[-2]) - 0.002) < 1e-9


def test_action_ids_bit_exact(golden_dir):
    g = _load(golden_dir, "action_ids.npz")
    enc = O.action_ids_idx_encode(torch.from_numpy(g["actions"]))
    assert (enc.numpy() == g["encoded"]).all()
    assert (enc.to(torch.float16).float().numpy() == g["encoded_fp16"]).all()


def test_unet_forward_matches_reference(golden_dir, unet_sd):
    g = _load(golden_dir, "unet_tiny.npz")
    taps = {}
    with torch.no_grad():
        out = O.unet_forward(unet_sd, UCFG, torch.from_numpy(g["sample"]), torch.tensor(float(g["timestep"])),
                             torch.from_numpy(g["enc"]), torch.from_numpy(g["added_time_ids"]),
                             torch.from_numpy(g["added_action_ids"]), taps=taps)
    for k in ("down.0.0", "mid", "up.3.2"):
        np.testing.assert_allclose(taps[k].numpy(), g["tap." + k], rtol=0, atol=2e-4)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-4)


def test_vae_decode_matches_reference(golden_dir, vae_sd):
    g = _load(golden_dir, "vae_tiny.npz")
    with torch.no_grad():
        out = O.vae_decode(vae_sd, VCFG, torch.from_numpy(g["z"]), TINY_UNET.num_frames)
    np.testing.assert_allclose(out.numpy(), g["out"], rtol=0, atol=2e-4)


def test_rollout_matches_reference_pipeline(golden_dir, unet_sd, vae_sd):
    """Whole vendored StableVideoDiffusionPipeline.__call__ (4 steps, CFG) vs the restated loop."""
    g = _load(golden_dir, "rollout_tiny.npz")
    n = int(g["num_steps"])
    sched = O.EulerSchedule(n)
    lat = torch.from_numpy(g["latents_raw"]) * sched.init_noise_sigma
    il = torch.from_numpy(g["image_latents"])
    ie = torch.from_numpy(g["image_embeddings"])
    assert float(il[0].abs().max()) == 0.0 and float(ie[0].abs().max()) == 0.0    # CFG rows are zeros
    with torch.no_grad():
        final = O.denoise_rollout(unet_sd, UCFG, lat, il[1:], ie[1:], torch.from_numpy(g["added_action_ids"]), n)
        frames = O.decode_and_postprocess(vae_sd, VCFG, final)
    np.testing.assert_allclose(final.numpy(), g["final_latents"], rtol=0, atol=5e-4)
    assert O.psnr(frames.numpy(), g["frames"]) > 70.0


def test_latent_noise_roll():
    """sample_latent_noise: action 2 / 3 roll the previous frame by +-w/16 (pipeline :750-823)."""
    noise = torch.arange(1 * 4 * 1 * 1 * 16, dtype=torch.float32).reshape(1, 4, 1, 1, 16)
    acts = torch.tensor([[4, 2, 3, 1]])
    out = O.sample_latent_noise(noise, acts)
    assert torch.equal(out[0, 1], torch.roll(noise[0, 0], 1, -1))
    assert torch.equal(out[0, 2], torch.roll(out[0, 1], -1, -1))
    assert torch.equal(out[0, 3], noise[0, 3])
    with pytest.raises(AssertionError):
        O.sample_latent_noise(torch.zeros(1, 2, 1, 1, 8), torch.tensor([[4, 2]]))
