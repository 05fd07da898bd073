"""Generate tests/golden/rollout_northstar.npz and tests/golden/latent_noise.npz from the reference's OWN code.

rollout_northstar.npz -- the complete vendored StableVideoDiffusionPipeline.__call__ at the BASELINE configs[1] geometry:
    576x1024 px (72x128 latents), 14 frames, classifier-free guidance, full 1.53 B-parameter UNet + full temporal VAE,
    fp32 on the CPU, 3 Euler steps (each forward is 89.6 TFLOP, the decode 97 TFLOP: ~20 min here).  Frozen:
    the captured conditioning (image embedding, image latents), the final latents (fp32, the decoder's exact input) and
    the decoded frames as two strided sub-samples in fp16 (phase (0,0) stride 4 and phase (1,3) stride 8 -- a decoder that
    misplaces pixels cannot match both), plus per-frame means of the full-resolution frames.
    Serves two GPU tests: the full-size rollout and the full-size temporal-VAE decode (a17).

latent_noise.npz -- pipeline.prepare_latents / sample_latent_noise / rotate_by_degrees / decode_action_seq_frames
    (pipeline :325-362, :750-844) run with a seeded CPU generator on several action rows (a7).

Runs only in the build container (needs /root/reference).  Weights and inputs come from numpy PCG64 seeds so the GPU
box regenerates them bit-identically.  Test infrastructure.

    python oracle/make_golden_northstar_rollout.py [--noise-only]
"""
from __future__ import annotations

import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG  # noqa: E402  (sys.path for the vendored fork + the transformers-5 shim)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from wiw_b200 import weights as W  # noqa: E402
from wiw_b200.config import FULL_UNET, FULL_VAE  # noqa: E402
from wiw_b200.frontend import vae_encoder_param_spec  # noqa: E402

SEED_UNET, SEED_VAE = 41, 42
F_, H_PX, W_PX, N_STEPS = 14, 576, 1024, 3
ACTIONS = [[4, 1, 2, 2, 1, 3, 1, 1, 2, 1, 1, 3, 3, 1]]


@torch.no_grad()
def latent_noise() -> None:
    from diffusers.pipelines.stable_video_diffusion import pipeline_stable_video_diffusion as P
    rows = torch.tensor([[4, 1, 2, 2, 1, 3, 1, 1, 2, 1, 1, 3, 3, 1],
                         [4, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2],
                         [4, 3, 1, 3, 1, 3, 2, 3, 2, 1, 1, 1, 0, 0]], dtype=torch.float32)
    aids = MG.action_ids_idx_encode(rows.shape[0], rows)
    decoded = P.decode_action_seq_frames(aids)
    d = {"actions": rows.numpy().astype(np.int64), "decoded": decoded.numpy()}
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        shape = (rows.shape[0], F_, 4, 8, 32)                       # w = 32 -> shift 2 columns
        g = torch.Generator().manual_seed(1234)
        d[f"noise_{name}"] = P.sample_latent_noise(decoded, shape, torch.device("cpu"), dt, g).float().numpy()
        d[f"seed_{name}"] = np.int64(1234)
    # prepare_latents end to end (the scheduler's init_noise_sigma multiply, pipeline :359-361)
    class _U:                                                       # the two attributes prepare_latents reads
        task_type = "navigation"
    class _Self:
        vae_scale_factor = 8
        unet = _U()
        scheduler = MG.svd_scheduler()
    _Self.scheduler.set_timesteps(25)
    g = torch.Generator().manual_seed(77)
    lat = P.StableVideoDiffusionPipeline.prepare_latents(_Self(), 3, F_, 8, 64, 256, torch.float16, torch.device("cpu"), g, None, aids)
    d["prepared_f16"] = lat.float().numpy()
    d["prepared_seed"] = np.int64(77)
    np.savez_compressed(os.path.join(MG.OUT, "latent_noise.npz"), **d)
    print("latent_noise.npz written", {k: v.shape for k, v in d.items() if hasattr(v, "shape")}, flush=True)


@torch.no_grad()
def rollout() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    c = FULL_UNET
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
        cap["final_latents"] = latents.clone()
        print(f"denoising done at {time.time() - t1:.0f} s", flush=True)
        return _dl(latents, *a, **k)

    pipe._encode_image, pipe._encode_vae_image, pipe.decode_latents = ei, ev, dl
    acts = torch.tensor(ACTIONS, dtype=torch.float32)
    aids = MG.action_ids_idx_encode(1, acts).float()
    img = Image.fromarray(np.random.Generator(np.random.PCG64(411)).integers(0, 256, (H_PX, W_PX, 3), dtype=np.uint8))
    lat_raw = MG.rng_normal(412, (1, F_, 4, H_PX // 8, W_PX // 8))
    t1 = time.time()
    frames = pipe(img, generator=torch.Generator().manual_seed(1), output_type="pt", height=H_PX, width=W_PX, num_frames=F_,
                  fps=7, motion_bucket_id=127, noise_aug_strength=0.02, num_inference_steps=N_STEPS, added_action_ids=aids,
                  latents=lat_raw.clone(), decode_chunk_size=F_).frames          # [1, F, 3, H, W] in [0, 1]
    print(f"pipeline ran in {time.time() - t1:.0f} s; frames {tuple(frames.shape)} mean {float(frames.mean()):.4f} "
          f"std {float(frames.std()):.4f}; final latents std {float(cap['final_latents'].std()):.4f}", flush=True)
    fr = frames[0]
    np.savez_compressed(os.path.join(MG.OUT, "rollout_northstar.npz"), num_steps=np.int64(N_STEPS),
                        seed_unet=np.int64(SEED_UNET), seed_vae=np.int64(SEED_VAE), seed_latents=np.int64(412),
                        actions=acts.numpy().astype(np.int64),
                        image_embeddings=cap["image_embeddings"].numpy(), image_latents=cap["image_latents"].numpy(),
                        added_action_ids=aids.numpy(), final_latents=cap["final_latents"].numpy(),
                        frames_s4=fr[:, :, ::4, ::4].numpy().astype(np.float16),
                        frames_s8_p13=fr[:, :, 1::8, 3::8].numpy().astype(np.float16),
                        frame_means=fr.mean(dim=(-1, -2)).numpy(), frame_sq_means=fr.pow(2).mean(dim=(-1, -2)).numpy())


if __name__ == "__main__":
    latent_noise()
    if "--noise-only" not in sys.argv:
        rollout()
