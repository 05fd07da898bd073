"""Rollout driver: the hot loop of StableVideoDiffusionPipeline.__call__ on the sm_100a kernels.

Reference: FTsvd/diffusers-private/diffusers/pipelines/stable_video_diffusion/
pipeline_stable_video_diffusion.py -- prepare_latents :325-362 (+ sample_latent_noise :750-786),
guidance scale :576-581, denoising loop :587-622, decode_latents :282-309, postprocess.
One `Rollout` serves one planner branch (request batch B == 1 with classifier-free guidance,
i.e. nb = 2 UNet samples), which is how the WM manager feeds its workers
(downstream/utils/worker_manager.py:448-469, --batch_size 1) and how branches shard over GPUs.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import lib as L
from . import ops
from .config import UNetConfig, VAEConfig
from .engine import UNetEngine, VAEDecoderEngine, sinusoid
from .scheduler import EulerSchedule


def sample_latent_noise_(noise: torch.Tensor, actions: torch.Tensor) -> torch.Tensor:
    """In-place action-correlated initial noise (pipeline :750-823): frame f >= 1 with action 2 (turn left)
    / 3 (turn right) becomes roll(frame f-1, +w/16 / -w/16) along width.  noise [F, 4, h, w]; integer semantics."""
    w = noise.shape[-1]
    if w % 16 != 0:
        raise AssertionError("Width must be divisible by 16 for 22.5-degree rotation.")
    shift = w // 16
    for i in range(1, noise.shape[0]):
        a = int(actions[i])
        if a == 2:
            noise[i] = torch.roll(noise[i - 1].clone(), shifts=shift, dims=-1)
        elif a == 3:
            noise[i] = torch.roll(noise[i - 1].clone(), shifts=-shift, dims=-1)
    return noise


class Rollout:
    def __init__(self, unet: UNetEngine, vae: Optional[VAEDecoderEngine], device="cuda", use_cuda_graph: bool = True):
        self.unet, self.vae = unet, vae
        self.use_cuda_graph = use_cuda_graph
        self.device = torch.device(device)
        self.F = unet.cfg.num_frames
        self._sched_cache = {}

    def _schedule(self, n: int):
        if n not in self._sched_cache:
            s = EulerSchedule(n)
            c0 = self.unet.cfg.block_out_channels[0]
            t_emb = sinusoid(s.timesteps, c0).to(torch.float16).to(self.device).contiguous()      # [N, C0]
            self._sched_cache[n] = (s, s.all_steps(), t_emb)
        return self._sched_cache[n]

    @torch.no_grad()
    def denoise(self, latents: torch.Tensor, image_latents: torch.Tensor, image_embedding: torch.Tensor,
                action_ids: torch.Tensor, num_steps: int = 25, noise_aug_strength: float = 0.02,
                min_guidance: float = 1.0, max_guidance: float = 3.0, trace: Optional[list] = None,
                cfg_split_group=None) -> torch.Tensor:
        """latents fp16 [F, 4, h, w] ALREADY scaled by init_noise_sigma (pipeline :359-361);
        image_latents [4, h, w]; image_embedding [1, D]; action_ids [F, A].
        Returns the final latents, fp16 channels-last [F, h, w, 4].

        cfg_split_group: a 2-rank process group (intra-request parallelism when there are fewer
        branches than GPUs).  The two rows of the classifier-free-guidance pair (pipeline :589-603) are independent UNet
        samples: rank 0 of the group runs the unconditional row, rank 1 the conditional row (UNet batch 1 each), ONE
        all-gather of v per Euler step ([F,h,w,32] fp16, 8 MB at 72x128) rebuilds the pair, and both ranks apply the same
        (bit-exact) guidance + Euler update, so their latents stay identical without any further exchange."""
        lib = L.require_device()
        F = self.F
        _, _, h, w = latents.shape
        hw = h * w
        dev = self.device
        sched, steps, t_embs = self._schedule(num_steps)
        split_rank = -1
        if cfg_split_group is not None:
            import torch.distributed as dist
            if dist.get_world_size(cfg_split_group) != 2:
                raise ValueError("cfg_split_group must have exactly 2 ranks (unconditional row, conditional row)")
            split_rank = dist.get_rank(cfg_split_group)
            emb = image_embedding if split_rank == 1 else torch.zeros_like(image_embedding)      # CFG row 0 = zeros (pipeline :221-227)
            self.unet.set_condition(emb, noise_aug_strength, action_ids, cfg_rows=False)
        else:
            self.unet.set_condition(image_embedding, noise_aug_strength, action_ids)
        lat = ops.nchw_to_nhwc(latents.to(dev, torch.float16).contiguous(), 4)                     # [F, h, w, 4]
        img = ops.nchw_to_nhwc(image_latents.to(dev, torch.float16).contiguous()[None], 4)[0]      # [h, w, 4]
        gs = torch.linspace(min_guidance, max_guidance, F).to(torch.float16).to(dev)               # pipeline :576-577
        x_in = torch.zeros(2 * F, h, w, 64, dtype=torch.float16, device=dev)
        st = L.stream_ptr
        v_pair = None
        for i, sc in enumerate(steps):
            L.check(lib.wiw_sampler_prepare(lat.data_ptr(), img.data_ptr(), x_in.data_ptr(), F, hw, sc.div, st()))
            if split_rank >= 0:
                import torch.distributed as dist
                v_half = (self.unet.forward_graphed if self.use_cuda_graph else self.unet.forward)(
                    x_in[split_rank * F:(split_rank + 1) * F], t_embs[i:i + 1])
                if v_pair is None:
                    v_pair = torch.empty((2 * F,) + tuple(v_half.shape[1:]), dtype=v_half.dtype, device=dev)
                dist.all_gather_into_tensor(v_pair, v_half.contiguous(), group=cfg_split_group)
                v = v_pair
            else:
                v = (self.unet.forward_graphed if self.use_cuda_graph else self.unet.forward)(x_in, t_embs[i:i + 1])
            L.check(lib.wiw_sampler_step(v.data_ptr(), lat.data_ptr(), gs.data_ptr(), F, hw, v.shape[-1], sc.c_out,
                                         sc.sig2p1, sc.sigma, sc.dt, st()))
            if trace is not None:
                trace.append(lat.clone())
        return lat

    @torch.no_grad()
    def decode(self, lat: torch.Tensor, want_f32: bool = False):
        """decode_latents + postprocess: lat fp16 [F, h, w, 4] -> uint8 [F, 3, H, W] (and optionally the
        un-quantised decoder output fp32 [F, 3, H, W] in [-1, 1])."""
        assert self.vae is not None
        F, h, w, _ = lat.shape
        z = torch.zeros(F, h, w, 64, dtype=torch.float16, device=self.device)
        # `1 / scaling_factor * latents` (pipeline :284): python float times fp16 tensor -> fp16
        z[..., :4] = (1 / self.vae.cfg.scaling_factor) * lat
        H, W = h * 8, w * 8
        u8 = torch.empty(F, 3, H, W, dtype=torch.uint8, device=self.device)
        f32 = torch.empty(F, 3, H, W, dtype=torch.float32, device=self.device) if want_f32 else None
        self.vae.decode(z, F, out_u8=u8, out_f32=f32)
        return (u8, f32) if want_f32 else u8
