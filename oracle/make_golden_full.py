"""Generate tests/golden/rollout_cfg1_full.npz: BASELINE.json configs[0] on the reference's OWN code.

    "FTsvd/eval_inference.py single 256x256 8-frame 10-step SVD rollout on CPU, random-init UNet, fixed seed"

Runs only in the build container (needs /root/reference, read-only; ~10 min of CPU).  The complete
vendored StableVideoDiffusionPipeline.__call__ is driven with the FULL architecture --
UNetSpatioTemporalConditionModel (320,640,1280,1280), heads (5,10,20,20), 1.53 B parameters, and the
temporal VAE (128,256,512,512) x 2 layers -- at 256x256 px (32x32 latents), F = 8, 10 Euler steps with
classifier-free guidance, fp32 on the CPU.  Weights are wiw_b200.weights.random_init (numpy PCG64:
the GPU box regenerates them bit-identically), so the fixture holds only the conditioning the pipeline
computed (CLIP projection, VAE-encoded image latents), the injected noise, the final latents and the
decoded frames.  Test infrastructure only -- nothing here is the product.

    python oracle/make_golden_full.py
"""
from __future__ import annotations

import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG          # sets sys.path for the vendored fork + the transformers-5 shim  # noqa: E402

import numpy as np                # noqa: E402
import torch                      # noqa: E402

from wiw_b200 import weights as W                        # noqa: E402
from wiw_b200.config import FULL_VAE, UNetConfig         # noqa: E402
from wiw_b200.frontend import vae_encoder_param_spec     # noqa: E402

SEED_UNET, SEED_VAE = 21, 22
F_, H_PX, W_PX, N_STEPS = 8, 256, 256, 10
CFG1_UNET = UNetConfig(num_frames=F_, action_input_channel=F_)      # full channels / heads, 8 frames (configs[0])


@torch.no_grad()
def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    c = CFG1_UNET
    unet = MG.UNetSpatioTemporalConditionModel(
        block_out_channels=c.block_out_channels, num_attention_heads=c.num_attention_heads,
        cross_attention_dim=c.cross_attention_dim, action_strategy="micro_cond", task_type="navigation",
        num_frames=c.num_frames, action_input_channel=c.action_input_channel)
    unet.load_state_dict(W.random_init(W.unet_param_spec(c), SEED_UNET), strict=True)
    unet.eval()
    v = FULL_VAE
    torch.manual_seed(0)
    vae = MG.AutoencoderKLTemporalDecoder(block_out_channels=v.block_out_channels, layers_per_block=v.layers_per_block,
                                          latent_channels=v.latent_channels, down_block_types=("DownEncoderBlock2D",) * 4)
    vae.decoder.load_state_dict(W.random_init(W.vae_decoder_param_spec(v), SEED_VAE), strict=True)
    esd = W.random_init(vae_encoder_param_spec(v.block_out_channels, v.layers_per_block, v.latent_channels), SEED_VAE + 1)
    missing, unexpected = vae.load_state_dict(esd, strict=False)
    assert not unexpected and all(k.startswith("decoder.") for k in missing)
    vae.eval()
    print(f"modules built in {time.time() - t0:.0f} s", flush=True)

    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModelWithProjection
    from PIL import Image
    torch.manual_seed(1)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=2, image_size=224,
        patch_size=32, projection_dim=c.cross_attention_dim)).eval()
    pipe = MG.StableVideoDiffusionPipeline(vae=vae, image_encoder=clip, unet=unet, scheduler=MG.svd_scheduler(),
                                           feature_extractor=CLIPImageProcessor())
    pipe.set_progress_bar_config(disable=True)
    cap = {}
    _ei, _ev, _dl = pipe._encode_image, pipe._encode_vae_image, pipe.decode_latents

    def ei(*a, **k):
        r = _ei(*a, **k); cap["image_embeddings"] = r.clone(); return r

    def ev(*a, **k):
        r = _ev(*a, **k); cap["image_latents"] = r.clone(); return r

    def dl(latents, *a, **k):
        cap["final_latents"] = latents.clone(); return _dl(latents, *a, **k)

    pipe._encode_image, pipe._encode_vae_image, pipe.decode_latents = ei, ev, dl
    acts = torch.tensor([[4, 1, 2, 2, 1, 3, 1, 1]], dtype=torch.float32)
    aids = MG.action_ids_idx_encode(1, acts).float()
    img = Image.fromarray(np.random.Generator(np.random.PCG64(311)).integers(0, 256, (H_PX, W_PX, 3), dtype=np.uint8))
    lat_raw = MG.rng_normal(312, (1, F_, 4, H_PX // 8, W_PX // 8))
    t1 = time.time()
    frames = pipe(img, generator=torch.Generator().manual_seed(1), output_type="pt", height=H_PX, width=W_PX, num_frames=F_,
                  fps=7, motion_bucket_id=127, noise_aug_strength=0.02, num_inference_steps=N_STEPS, added_action_ids=aids,
                  latents=lat_raw.clone(), decode_chunk_size=F_).frames
    print(f"pipeline ran in {time.time() - t1:.0f} s; frames {tuple(frames.shape)} mean {float(frames.mean()):.4f} "
          f"std {float(frames.std()):.4f}; final latents std {float(cap['final_latents'].std()):.4f}", flush=True)
    np.savez_compressed(os.path.join(MG.OUT, "rollout_cfg1_full.npz"), latents_raw=lat_raw.numpy(), num_steps=np.int64(N_STEPS),
                        seed_unet=np.int64(SEED_UNET), seed_vae=np.int64(SEED_VAE), actions=acts.numpy().astype(np.int64),
                        image_embeddings=cap["image_embeddings"].numpy(), image_latents=cap["image_latents"].numpy(),
                        added_action_ids=aids.numpy(), final_latents=cap["final_latents"].numpy(),
                        frames_f16=frames.numpy().astype(np.float16))


if __name__ == "__main__":
    main()
