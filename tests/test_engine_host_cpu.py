"""Launch sequencing of the product engines checked WITHOUT a GPU: `UNetEngine.forward`, `Rollout.denoise` and
`VAEDecoderEngine.decode` run on CPU tensors with every C-ABI call served by the host emulator of include/wiw.h
(tests/helpers/cabi_emulator.py -- test infrastructure, fp32 arithmetic) and are compared with the golden outputs of the
reference's vendored fork.  What this pins on the CPU: the order and operands of ~750 launches per forward (residual /
AlphaBlender wiring, the collapsed cross-attention vectors, the fused time-embedding GEMM and its column offsets, the frame-position
tables, skip concatenation, up / down-sampling), the sampler loop, and the decode tail.  The kernels' own arithmetic is what
the `-m gpu` tests check."""
import os
import sys

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

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import cabi_emulator  # noqa: E402


@pytest.fixture()
def emu(monkeypatch):
    return cabi_emulator.install(monkeypatch)


def rel_rms(a, b):
    return float(((a - b).pow(2).mean() / b.pow(2).mean()).sqrt())


def test_unet_forward_sequencing_vs_golden(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "unet_tiny.npz"))
    sample, enc, aids = torch.from_numpy(g["sample"]), torch.from_numpy(g["enc"]), torch.from_numpy(g["added_action_ids"])
    nb, F, _, h, w = sample.shape
    unet = UNetEngine(TINY_UNET, W.random_init(W.unet_param_spec(TINY_UNET), 11), "cpu")
    unet.set_condition(enc[1], 0.02, aids[0])
    x_in = ops.nchw_to_nhwc(sample.flatten(0, 1).half().contiguous(), 64)
    t_emb = sinusoid(torch.tensor([float(g["timestep"])]), TINY_UNET.block_out_channels[0]).half()
    taps = {}
    v = unet.forward(x_in, t_emb, taps=taps)
    out = ops.nhwc_to_nchw(v, 4).float().view(nb, F, 4, h, w)
    for k in ("down.0.0", "mid", "up.3.2"):
        got = ops.nhwc_to_nchw(taps[k], taps[k].shape[-1]).float()
        assert rel_rms(got, torch.from_numpy(g["tap." + k])) < 1e-2, k
    assert rel_rms(out, torch.from_numpy(g["out"])) < 1e-2
    n_gemm = len(emu.calls)
    assert n_gemm > 250                                    # every Linear / conv of the forward went through the descriptor path


def test_rollout_and_decode_sequencing_vs_reference_pipeline(emu, golden_dir):
    g = np.load(os.path.join(golden_dir, "rollout_tiny.npz"))
    n = int(g["num_steps"])
    unet = UNetEngine(TINY_UNET, W.random_init(W.unet_param_spec(TINY_UNET), 11), "cpu")
    vae = VAEDecoderEngine(TINY_VAE, W.random_init(W.vae_decoder_param_spec(TINY_VAE), 12), "cpu")
    ro = Rollout(unet, vae, "cpu", use_cuda_graph=False)
    lat0 = (torch.from_numpy(g["latents_raw"]).half() * EulerSchedule(n).init_noise_sigma)[0]
    lat = ro.denoise(lat0, torch.from_numpy(g["image_latents"])[1], torch.from_numpy(g["image_embeddings"])[1],
                     torch.from_numpy(g["added_action_ids"])[0], num_steps=n)
    final = ops.nhwc_to_nchw(lat, 4).float()
    ref_lat = torch.from_numpy(g["final_latents"])[0]
    rng = float(ref_lat.max() - ref_lat.min())
    psnr_lat = 20 * np.log10(rng / float((final - ref_lat).pow(2).mean().sqrt()))
    assert psnr_lat > 45.0, psnr_lat
    u8, f32 = ro.decode(lat, want_f32=True)
    frames = (f32 / 2 + 0.5).clamp(0, 1).numpy()
    assert O.psnr(frames, g["frames"][0]) >= 45.0
    assert O.psnr(u8.numpy().astype(np.float32) / 255.0, g["frames"][0]) >= 45.0


def test_vae_encoder_sequencing_vs_golden(emu, golden_dir):
    """the SD-VAE encoder + quant_conv on the implicit-GEMM / GroupNorm wrappers reproduces the
    reference's fp32 posterior mean (frontend_tiny.npz) -- asymmetric stride-2 taps, shortcut convs, single-head attention and
    the quant_conv fold."""
    from wiw_b200.engine import VAEEncoderEngine
    from wiw_b200.frontend import vae_encoder_param_spec
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    c = TINY_VAE
    esd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), 13)
    enc = VAEEncoderEngine(c, esd, "cpu")
    got = enc.encode_mode(torch.from_numpy(g["image"])).float()
    ref = torch.from_numpy(g["latent_mode"])
    assert got.shape == ref.shape
    assert rel_rms(got, ref) < 1e-2, rel_rms(got, ref)
