"""CLIP image encoder on the sm_100a kernels (wiw_b200/clip.py) against the installed transformers module on the same weights --
the third-party arithmetic the reference calls in _encode_image (pipeline :214), absent from /root/reference (parity anchored on
the installed package, as SURVEY.md section 8c says for third-party dependencies).

Tolerance: fp16 storage / fp32 accumulation against the fp32 module: relative RMS <= 2e-2 on image_embeds, cosine >= 0.999."""
import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import clip as C
from wiw_b200 import lib as L

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _dense(n_in, n_out):
    b, w = C.antialias_operator(n_in, n_out)
    m = np.zeros((n_out, n_in))
    for o in range(n_out):
        m[o, b[o, 0]: b[o, 0] + b[o, 1]] = w[o, : b[o, 1]]
    return m


@pytest.mark.parametrize("H,W", [(480, 480), (384, 384), (576, 1024), (61, 47)])
def test_preprocessor_kernel_matches_pipeline_resize(H, W):
    """uint8 observation -> blur + bicubic(align_corners) + CLIP normalisation in one kernel pair, against the pipeline's own
    _resize_with_antialiasing formula (oracle/frontend_oracle.py, pinned on reference goldens) in fp32."""
    from oracle.frontend_oracle import CLIP_MEAN, CLIP_STD, resize_with_antialiasing
    img = torch.from_numpy(np.random.default_rng(H + W).integers(0, 256, size=(2, 3, H, W), dtype=np.uint8))
    got = C.ClipPreprocessor(DEV)(img.to(DEV)).float().cpu()
    x = (resize_with_antialiasing(img.float() / 255.0 * 2.0 - 1.0, (224, 224)) + 1.0) / 2.0
    want = (x - torch.tensor(CLIP_MEAN).view(1, 3, 1, 1)) / torch.tensor(CLIP_STD).view(1, 3, 1, 1)
    assert got.shape == want.shape == (2, 3, 224, 224)
    assert float((got - want).abs().max()) < 4e-3            # fp16 output rounding of values up to ~2.7


def test_attention_small_matches_sdpa():
    lib = L.require_device()
    for (B, S, heads, d) in [(2, 257, 16, 80), (1, 50, 2, 32), (3, 77, 4, 64)]:
        g = torch.Generator().manual_seed(S)
        qkv = torch.randn(B * S, 3 * heads * d, generator=g).half().to(DEV)
        out = torch.empty(B * S, heads * d, dtype=torch.float16, device=DEV)
        L.check(lib.wiw_attention_small(qkv.data_ptr(), out.data_ptr(), B, S, heads, d, d ** -0.5, L.stream_ptr()))
        t = qkv.float().view(B, S, 3, heads, d)
        q, k, v = (t[:, :, i].permute(0, 2, 1, 3) for i in range(3))
        ref = torch.nn.functional.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * S, heads * d)
        assert float((out.float() - ref).abs().max()) < 3e-3, (B, S, heads, d)


@pytest.mark.parametrize("name,hidden,inter,layers,heads,patch,act,proj", [
    ("tiny-quick", 128, 256, 2, 2, 32, "quick_gelu", 64),
    ("d80-gelu", 320, 640, 2, 4, 32, "gelu", 64),
    ("vit-h-14", 1280, 5120, 32, 16, 14, "gelu", 1024),
])
def test_clip_engine_vs_transformers(name, hidden, inter, layers, heads, patch, act, proj):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(0)
    cfg = CLIPVisionConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=heads, image_size=224,
                           patch_size=patch, projection_dim=proj, hidden_act=act)
    ref = CLIPVisionModelWithProjection(cfg).eval()
    eng = C.ClipVisionEngine.from_module(ref, DEV)
    img = torch.from_numpy(np.random.default_rng(3).integers(0, 256, size=(2, 3, 480, 480), dtype=np.uint8)).to(DEV)
    pixels = eng.pre(img)                                     # same (fp16) pixels to both: isolates the encoder
    got = eng.encode_pixels(pixels).float().cpu()
    with torch.no_grad():
        want = ref.to(DEV)(pixels.float()).image_embeds.float().cpu()
    torch.cuda.synchronize()
    assert got.shape == want.shape == (2, proj) and torch.isfinite(got).all()
    rel = float(((got - want).pow(2).mean() / want.pow(2).mean()).sqrt())
    cos = float(torch.nn.functional.cosine_similarity(got, want, dim=-1).min())
    from conftest import record_parity
    record_parity(f"clip_engine_{name}_vs_transformers_fp32", rel_rms=rel, min_cosine=cos)
    assert rel < 2e-2 and cos > 0.999, (rel, cos)
