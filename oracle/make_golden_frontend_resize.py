"""Generate tests/golden/frontend_resize.npz: the reference pipeline's OWN conditioning front-end on an observation that is NOT
already height x width (the solvers send 384x384 / 480x480 frames, the pipeline runs at 576x1024).

Calls, unmodified, StableVideoDiffusionPipeline._encode_image (pipeline :183-229: original-size image -> _resize_with_antialiasing ->
224x224 -> CLIP), video_processor.preprocess(image, height, width) (PIL LANCZOS resize in uint8 + normalise; DP/image_processor.py:501)
and _encode_vae_image (:231-252) on a 48x40 RGB image with the tiny VAE (64x128 px target) and a tiny CLIP whose weights are stored
in the fixture (third-party `transformers` arithmetic, re-instantiated on the GPU box from these weights).  noise_aug_strength = 0 so
the frozen latents do not depend on a generator.  Runs only in the build container.  Test infrastructure.

    python oracle/make_golden_frontend_resize.py
"""
from __future__ import annotations

import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG  # noqa: E402

import numpy as np  # noqa: E402
import torch  # noqa: E402

CLIP_CFG = dict(hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, image_size=224, patch_size=32)


@torch.no_grad()
def main() -> None:
    from PIL import Image
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModelWithProjection
    c = MG.TINY_UNET
    vae = MG.build_vae()
    torch.manual_seed(1)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(projection_dim=c.cross_attention_dim, **CLIP_CFG)).eval()
    pipe = MG.StableVideoDiffusionPipeline(vae=vae, image_encoder=clip, unet=MG.build_unet(), scheduler=MG.svd_scheduler(),
                                           feature_extractor=CLIPImageProcessor())
    out = {}
    for tag, (hin, win) in (("small", (40, 48)), ("square", (96, 96))):
        img_u8 = np.random.Generator(np.random.PCG64(501 + hin)).integers(0, 256, (hin, win, 3), dtype=np.uint8)
        img = [Image.fromarray(img_u8)]
        emb = pipe._encode_image(img, torch.device("cpu"), 1, True)                              # [2, 1, D], row 0 zeros
        pre = pipe.video_processor.preprocess(img, height=MG.H_PX, width=MG.W_PX)                # [1, 3, H, W] in [-1, 1]
        lat = pipe._encode_vae_image(pre, torch.device("cpu"), 1, True)                          # [2, 4, h, w], row 0 zeros
        out[f"{tag}_image"] = img_u8.transpose(2, 0, 1).copy()
        out[f"{tag}_image_embeddings"] = emb.numpy()
        out[f"{tag}_preprocessed_u8"] = np.round((pre[0].numpy() + 1.0) * 127.5).astype(np.uint8)   # exact: pre = u8 / 255 * 2 - 1
        out[f"{tag}_image_latents"] = lat.numpy()
    for k, v in clip.state_dict().items():
        out["clip." + k] = v.numpy()
    np.savez_compressed(os.path.join(MG.OUT, "frontend_resize.npz"), height=np.int64(MG.H_PX), width=np.int64(MG.W_PX),
                        seed_vae=np.int64(MG.SEED_VAE), **out)
    print({k: v.shape for k, v in out.items() if not k.startswith("clip.")})


if __name__ == "__main__":
    main()
