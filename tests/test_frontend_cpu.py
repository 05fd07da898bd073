"""CPU parity of the conditioning front-end (wiw_b200/frontend.py, torch ops) against the vendored
fork's outputs in tests/golden/frontend_tiny.npz."""
import os

import numpy as np
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import weights as W
from wiw_b200.config import TINY_VAE
from wiw_b200.frontend import vae_encode_mode, vae_encoder_param_spec
from oracle.frontend_oracle import resize_with_antialiasing


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


def _tiny_clip(g, dtype=torch.float32, device="cpu"):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    sd = {k[len("clip."):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("clip.")}
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2,
                                                          image_size=224, patch_size=32, projection_dim=sd["visual_projection.weight"].shape[0]))
    clip.load_state_dict(sd, strict=True)
    return clip.to(device, dtype).eval()


def test_observation_front_end_matches_reference_pipeline(golden_dir):
    """An observation that is NOT height x width (the solvers send 384x384 / 480x480): the CLIP branch sees the image as sent
    (pipeline :192-200), the VAE branch a PIL-LANCZOS uint8 resize (video_processor.preprocess).  Golden = the reference
    pipeline's own _encode_image / preprocess / _encode_vae_image (oracle/make_golden_frontend_resize.py)."""
    from wiw_b200 import resize as R
    from oracle.frontend_oracle import encode_image_clip
    g = np.load(os.path.join(golden_dir, "frontend_resize.npz"))
    H, Wd = int(g["height"]), int(g["width"])
    c = TINY_VAE
    sd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), int(g["seed_vae"]) + 1)
    clip = _tiny_clip(g)
    for tag in ("small", "square"):
        img = g[f"{tag}_image"]                                                         # uint8 [3, Hin, Win]
        pre_u8 = R.resize_reference(np.ascontiguousarray(img.transpose(1, 2, 0)), Wd, H, "lanczos").transpose(2, 0, 1)
        assert np.array_equal(pre_u8, g[f"{tag}_preprocessed_u8"]), tag               # byte-exact twin of the PIL LANCZOS step
        emb = encode_image_clip(torch.from_numpy(img)[None].float() / 255.0, clip)
        np.testing.assert_allclose(emb.numpy(), g[f"{tag}_image_embeddings"][1:], rtol=0, atol=2e-5)
        assert float(np.abs(g[f"{tag}_image_embeddings"][0]).max()) == 0.0              # CFG row
        pre = (torch.from_numpy(pre_u8)[None].float() / 255.0) * 2.0 - 1.0
        lat = vae_encode_mode(sd, pre, len(c.block_out_channels), c.layers_per_block)
        np.testing.assert_allclose(lat.numpy(), g[f"{tag}_image_latents"][1:], rtol=0, atol=2e-5)
