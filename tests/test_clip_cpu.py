"""Host logic of the CLIP image encoder (wiw_b200/clip.py) without a GPU: the composed antialias-resize operator against the
pipeline's own formula, and the launch sequencing of ClipVisionEngine (patch GEMM + position table, fused qkv, residual epilogues,
GELU / quick-GELU, pooling, projection) through the C-ABI emulator against the installed transformers module -- the third-party
arithmetic the reference itself calls (pipeline :214)."""
import os
import sys

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import clip as C

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
import cabi_emulator  # noqa: E402


@pytest.fixture
def emu(monkeypatch):
    return cabi_emulator.install(monkeypatch)


def _dense(n_in, n_out):
    b, w = C.antialias_operator(n_in, n_out)
    m = np.zeros((n_out, n_in))
    for o in range(n_out):
        m[o, b[o, 0]: b[o, 0] + b[o, 1]] = w[o, : b[o, 1]]
    return m


@pytest.mark.parametrize("H,W", [(480, 480), (384, 384), (576, 1024), (40, 48)])
def test_antialias_operator_matches_reference_golden(H, W, golden_dir):
    """One banded operator per axis == Gaussian blur (reflect) + bicubic(align_corners=True) of pipeline :643-747."""
    g = np.load(os.path.join(golden_dir, "frontend_tiny.npz"))
    if (H, W) == (40, 48):                                         # reference-generated fixture (288x512 -> 224x224, fp16 storage)
        x = g["resize_in"].astype(np.float64)[0]
        got = np.einsum("oh,chw,pw->cop", _dense(288, 224), x, _dense(512, 224))
        np.testing.assert_allclose(got, g["resize_out"].astype(np.float64)[0], rtol=0, atol=2e-3)
        x = g["image"].astype(np.float64)[0]                        # up-sampling case (64x128 -> 224x224)
        got = np.einsum("oh,chw,pw->cop", _dense(64, 224), x, _dense(128, 224))
        np.testing.assert_allclose(got, g["resize_up_out"].astype(np.float64)[0], rtol=0, atol=2e-3)
    m = _dense(H, 224)
    assert np.allclose(m.sum(1), 1.0, atol=1e-6)                   # both kernels sum to one: the *2-1 / (x+1)/2 round trip cancels
    assert C.antialias_operator(H, 224)[1].shape[1] <= 12


@pytest.mark.parametrize("hidden,heads,act", [(128, 2, "quick_gelu"), (320, 4, "gelu")])
def test_clip_engine_sequencing_vs_transformers(emu, hidden, heads, act):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=hidden, intermediate_size=2 * hidden, num_hidden_layers=2, num_attention_heads=heads, image_size=224,
                           patch_size=32, projection_dim=64, hidden_act=act)
    ref = CLIPVisionModelWithProjection(cfg).eval()
    eng = C.ClipVisionEngine.from_module(ref, "cpu")
    img = torch.from_numpy(np.random.default_rng(1).integers(0, 256, size=(2, 3, 96, 80), dtype=np.uint8))
    got = eng(img).float()
    # the reference pre-processing, evaluated with the dense operators (the pipeline's resize is checked above)
    x = np.einsum("oh,bchw,pw->bcop", _dense(96, 224), img.numpy() / 255.0, _dense(80, 224))
    x = (x - np.array(C.CLIP_MEAN)[None, :, None, None]) / np.array(C.CLIP_STD)[None, :, None, None]
    with torch.no_grad():
        want = ref(torch.from_numpy(x).float()).image_embeds[:, None]
    assert got.shape == want.shape == (2, 1, 64)
    rel = float(((got - want).pow(2).mean() / want.pow(2).mean()).sqrt())
    assert rel < 1e-2, rel                                          # fp16 activations between layers vs the fp32 module
