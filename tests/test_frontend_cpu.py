"""CPU parity of the conditioning front-end (wiw_b200/frontend.py, torch ops) against the vendored
fork's outputs in tests/golden/frontend_tiny.npz."""
import os

import numpy as np
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import weights as W
from wiw_b200.config import TINY_VAE
from wiw_b200.frontend import resize_with_antialiasing, vae_encode_mode, vae_encoder_param_spec


def test_vae_encode_mode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    c = TINY_VAE
    sd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), 13)
    out = vae_encode_mode(sd, torch.from_numpy(g["image"]), len(c.block_out_channels), c.layers_per_block)
    np.testing.assert_allclose(out.numpy(), g["latent_mode"], rtol=0, atol=2e-5)


def test_resize_with_antialiasing_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    down = resize_with_antialiasing(torch.from_numpy(g["resize_in"]).float(), (224, 224))
    np.testing.assert_allclose(down.numpy(), g["resize_out"].astype(np.float32), rtol=0, atol=2e-3)
    up = resize_with_antialiasing(torch.from_numpy(g["image"]), (224, 224))
    np.testing.assert_allclose(up.numpy(), g["resize_up_out"].astype(np.float32), rtol=0, atol=2e-3)
