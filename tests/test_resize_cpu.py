"""The numpy twin of the GPU resize (wiw_b200/resize.py) is byte-identical to PIL's Image.resize(BICUBIC); the
requantisation LUT equals the reference's to_tensor -> (clip*255).astype(uint8) chain."""
import numpy as np
import pytest
import torch
from PIL import Image

import wiw_b200  # noqa: F401
from wiw_b200 import resize as R


@pytest.mark.parametrize("H,W,ow,oh", [(576, 1024, 480, 480), (64, 128, 48, 40), (40, 30, 64, 50), (17, 23, 5, 7)])
def test_numpy_twin_matches_pil(H, W, ow, oh):
    img = np.random.default_rng(H * W).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BICUBIC))
    assert np.array_equal(R.resize_reference(img, ow, oh), ref)


@pytest.mark.parametrize("H,W,ow,oh", [(480, 480, 1024, 576), (384, 384, 1024, 576), (64, 128, 48, 40), (17, 23, 40, 31), (576, 1024, 1024, 576)])
def test_numpy_twin_matches_pil_lanczos(H, W, ow, oh):
    """The request side: video_processor.preprocess resizes the conditioning image with PIL LANCZOS (DP/image_processor.py:501)."""
    img = np.random.default_rng(H * W + 1).integers(0, 256, (H, W, 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((ow, oh), Image.LANCZOS))
    assert np.array_equal(R.resize_reference(img, ow, oh, "lanczos"), ref)


def test_requant_lut_matches_reference_chain():
    from torchvision.transforms.functional import to_tensor
    img = Image.fromarray(np.arange(256, dtype=np.uint8).reshape(16, 16))
    t = to_tensor(img)                                            # float32 / 255 (api_models/__init__.py:164)
    q = (np.clip(t.numpy(), 0, 1) * 255).astype(np.uint8)         # api_models/__init__.py:229
    assert np.array_equal(R.requant_lut(), q.reshape(-1))
    assert (R.requant_lut().astype(int) <= np.arange(256)).all()  # truncation never rounds up
