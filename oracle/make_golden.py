"""Generate tests/golden/*.npz by running the reference's OWN vendored diffusers fork.

Runs only in the build container (needs /root/reference, read-only).  It imports
/root/reference/FTsvd/diffusers-private (diffusers 0.31.0, the modified fork) with the
one-line transformers-5 shim, loads the seeded weights of wiw_b200.weights.random_init into
the reference modules (strict), drives them on seeded inputs and freezes their outputs.
Nothing here is the restatement: every number written comes out of reference code.

    python oracle/make_golden.py            # rewrites tests/golden/*.npz

Fixtures (all fp32 CPU, tiny geometry = wiw_b200.config.TINY_*):
  scheduler.npz      sigmas / timesteps / init_noise_sigma for N in {10,25,30}
  action_ids.npz     utils.svd_utils.action_ids_idx_encode on fixed action rows
  unet_tiny.npz      one UNet forward (inputs, output, a few block taps)
  vae_tiny.npz       one temporal-VAE decode
  frontend_tiny.npz  VAE encode posterior mode + _resize_with_antialiasing (pipeline :231-252, :643-667)
  rollout_tiny.npz   the full vendored StableVideoDiffusionPipeline.__call__ (4 Euler steps,
                     CFG, injected latents): captured conditioning + final latents + frames
"""
from __future__ import annotations

import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(REF, "FTsvd", "diffusers-private"))
sys.path.insert(0, REF)
sys.path.insert(0, REPO)

import numpy as np
import torch

import transformers.utils
transformers.utils.FLAX_WEIGHTS_NAME = "flax_model.msgpack"   # removed in transformers 5; DP/pipelines/pipeline_loading_utils.py:49

from diffusers.models.unets.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel
from diffusers.models.autoencoders.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
from diffusers.schedulers.scheduling_euler_discrete import EulerDiscreteScheduler
from diffusers.pipelines.stable_video_diffusion.pipeline_stable_video_diffusion import StableVideoDiffusionPipeline
from utils.svd_utils import action_ids_idx_encode, get_action_ids

import wiw_b200  # noqa: F401
from wiw_b200 import weights as W
from wiw_b200.config import TINY_UNET, TINY_VAE

OUT = os.path.join(REPO, "tests", "golden")
SEED_UNET, SEED_VAE = 11, 12
H_PX, W_PX = 64, 128            # latent 8 x 16


def rng_normal(seed: int, shape) -> torch.Tensor:
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).standard_normal(shape, dtype=np.float32))


def svd_scheduler() -> EulerDiscreteScheduler:
    """The scheduler config of the SVD checkpoint (external file; restated in SURVEY.md 8c)."""
    return EulerDiscreteScheduler(
        beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear", prediction_type="v_prediction",
        timestep_type="continuous", use_karras_sigmas=True, sigma_min=0.002, sigma_max=700.0,
        timestep_spacing="leading", steps_offset=1, interpolation_type="linear", num_train_timesteps=1000)


def build_unet():
    c = TINY_UNET
    u = UNetSpatioTemporalConditionModel(
        block_out_channels=c.block_out_channels, num_attention_heads=c.num_attention_heads,
        cross_attention_dim=c.cross_attention_dim, action_strategy="micro_cond", task_type="navigation",
        num_frames=c.num_frames, action_input_channel=c.action_input_channel)
    sd = W.random_init(W.unet_param_spec(c), SEED_UNET)
    u.load_state_dict(sd, strict=True)
    return u.eval()


def build_vae():
    c = TINY_VAE
    torch.manual_seed(0)      # encoder + quant_conv keep torch default init (not part of the restatement)
    v = AutoencoderKLTemporalDecoder(block_out_channels=c.block_out_channels, layers_per_block=c.layers_per_block,
                                     latent_channels=c.latent_channels, down_block_types=("DownEncoderBlock2D",) * 4)
    sd = W.random_init(W.vae_decoder_param_spec(c), SEED_VAE)
    v.decoder.load_state_dict(sd, strict=True)
    from wiw_b200.frontend import vae_encoder_param_spec
    esd = W.random_init(vae_encoder_param_spec(c.block_out_channels, c.layers_per_block, c.latent_channels), SEED_VAE + 1)
    missing, unexpected = v.load_state_dict(esd, strict=False)
    assert not unexpected and all(k.startswith("decoder.") for k in missing), (missing[:3], unexpected[:3])
    return v.eval()


@torch.no_grad()
def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)

    # ---- scheduler tables -------------------------------------------------------------
    d = {}
    for n in (10, 25, 30):
        s = svd_scheduler()
        s.set_timesteps(n)
        d[f"sigmas_{n}"] = s.sigmas.numpy()
        d[f"timesteps_{n}"] = s.timesteps.numpy()
        d[f"init_noise_sigma_{n}"] = np.float32(s.init_noise_sigma.item())
        # step-index arithmetic: index_for_timestep of every timestep (scheduler :554-566)
        d[f"index_{n}"] = np.array([s.index_for_timestep(t) for t in s.timesteps], dtype=np.int64)
    np.savez(os.path.join(OUT, "scheduler.npz"), **d)

    # ---- action ids ---------------------------------------------------------------------
    acts = torch.tensor([[4, 1, 2, 2, 1, 3, 1, 1], [1, 3, 3, 0, 0, 2, 1, 4]], dtype=torch.float32)
    np.savez(os.path.join(OUT, "action_ids.npz"), actions=acts.numpy().astype(np.int64),
             encoded=action_ids_idx_encode(2, acts).numpy(),
             encoded_fp16=get_action_ids(2, acts, "micro_cond", torch.float16).float().numpy())

    # ---- UNet forward ---------------------------------------------------------------------
    c = TINY_UNET
    F_, h, w = c.num_frames, H_PX // 8, W_PX // 8
    unet = build_unet()
    sample = rng_normal(101, (2, F_, 8, h, w))
    enc = rng_normal(102, (2, 1, c.cross_attention_dim))
    enc[0] = 0
    tids = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    aids = action_ids_idx_encode(1, acts[:1]).float()
    tstep = torch.tensor(1.2345)
    taps = {}
    hooks = [
        unet.down_blocks[0].resnets[0].register_forward_hook(lambda m, i, o: taps.__setitem__("res.down.0.0", o)),
        unet.down_blocks[0].attentions[0].register_forward_hook(lambda m, i, o: taps.__setitem__("down.0.0", o[0])),
        unet.mid_block.register_forward_hook(lambda m, i, o: taps.__setitem__("mid", o)),
        unet.up_blocks[3].attentions[2].register_forward_hook(lambda m, i, o: taps.__setitem__("up.3.2", o[0])),
    ]
    out = unet(sample, tstep, encoder_hidden_states=enc, added_time_ids=tids, return_dict=False, added_action_ids=aids)[0]
    for hk in hooks:
        hk.remove()
    np.savez(os.path.join(OUT, "unet_tiny.npz"), sample=sample.numpy(), enc=enc.numpy(), added_time_ids=tids.numpy(),
             added_action_ids=aids.numpy(), timestep=np.float32(tstep.item()), out=out.numpy(),
             **{"tap." + k: v.numpy() for k, v in taps.items()})
    print("unet_tiny out", tuple(out.shape), float(out.abs().mean()), float(out.std()))

    # ---- VAE decode -------------------------------------------------------------------------
    vae = build_vae()
    z = rng_normal(201, (F_, 4, h, w))
    dec = vae.decode(z, num_frames=F_).sample
    np.savez(os.path.join(OUT, "vae_tiny.npz"), z=z.numpy(), out=dec.numpy())
    print("vae_tiny out", tuple(dec.shape), float(dec.abs().mean()), float(dec.std()))

    # ---- conditioning front-end: fp32 VAE encode (posterior mode) + antialiased CLIP resize --------
    from diffusers.pipelines.stable_video_diffusion.pipeline_stable_video_diffusion import _resize_with_antialiasing
    img_t = rng_normal(401, (1, 3, H_PX, W_PX)).clamp(-1, 1)
    enc_mode = vae.encode(img_t).latent_dist.mode()
    big = rng_normal(402, (1, 3, 288, 512)).clamp(-1, 1)
    np.savez(os.path.join(OUT, "frontend_tiny.npz"), image=img_t.numpy(), latent_mode=enc_mode.numpy(),
             resize_in=big.numpy().astype(np.float16), resize_out=_resize_with_antialiasing(big.half().float(), (224, 224)).numpy().astype(np.float16),
             resize_up_out=_resize_with_antialiasing(img_t, (224, 224)).numpy().astype(np.float16))
    print("frontend_tiny latent", tuple(enc_mode.shape), float(enc_mode.std()))

    # ---- full pipeline rollout ----------------------------------------------------------------
    from transformers import CLIPImageProcessor, CLIPVisionConfig, CLIPVisionModelWithProjection
    from PIL import Image
    torch.manual_seed(1)
    clip = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=32, intermediate_size=64, num_hidden_layers=1, num_attention_heads=2, image_size=224,
        patch_size=32, projection_dim=c.cross_attention_dim)).eval()
    pipe = StableVideoDiffusionPipeline(vae=vae, image_encoder=clip, unet=unet, scheduler=svd_scheduler(),
                                        feature_extractor=CLIPImageProcessor())
    pipe.set_progress_bar_config(disable=True)
    cap = {}
    _ei, _ev = pipe._encode_image, pipe._encode_vae_image

    def ei(*a, **k):
        r = _ei(*a, **k); cap["image_embeddings"] = r.clone(); return r

    def ev(*a, **k):
        r = _ev(*a, **k); cap["image_latents"] = r.clone(); return r

    pipe._encode_image, pipe._encode_vae_image = ei, ev
    img = Image.fromarray(np.random.Generator(np.random.PCG64(301)).integers(0, 256, (H_PX, W_PX, 3), dtype=np.uint8))
    lat_raw = rng_normal(302, (1, F_, 4, h, w))
    nsteps = 4
    kw = dict(height=H_PX, width=W_PX, num_frames=F_, fps=7, motion_bucket_id=127, noise_aug_strength=0.02,
              num_inference_steps=nsteps, added_action_ids=aids, latents=lat_raw.clone())
    final_lat = pipe(img, generator=torch.Generator().manual_seed(1), output_type="latent", **kw).frames
    frames = pipe(img, generator=torch.Generator().manual_seed(1), output_type="pt", **kw).frames
    np.savez(os.path.join(OUT, "rollout_tiny.npz"), latents_raw=lat_raw.numpy(), num_steps=np.int64(nsteps),
             image_embeddings=cap["image_embeddings"].numpy(), image_latents=cap["image_latents"].numpy(),
             added_action_ids=aids.numpy(), final_latents=final_lat.numpy(), frames=frames.numpy())
    print("rollout_tiny frames", tuple(frames.shape), float(frames.mean()), float(frames.std()))


if __name__ == "__main__":
    main()
