"""EDM / Euler sampler tables for the SVD-dagger rollout (host side, bit-exact with the reference).

Restates EulerDiscreteScheduler with the SVD checkpoint's config
(FTsvd/diffusers-private/diffusers/schedulers/scheduling_euler_discrete.py): v_prediction,
continuous timesteps, Karras sigmas (sigma_min 0.002, sigma_max 700, rho 7), 'leading' spacing.
The tables are produced with the same numpy-float64 / torch-float32 operations the reference
uses (set_timesteps :319-450, _convert_to_karras :476-499) so sigmas, timesteps and the step-index
arithmetic are bit-identical; the per-element update itself runs in wiw_sampler_step.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np
import torch


@dataclass
class StepScalars:
    """fp32 scalars of one Euler step, computed with the reference's own 0-dim tensor ops."""
    sigma: float
    div: float        # sqrt(sigma^2 + 1)          scale_model_input :313-314
    c_out: float      # -sigma / sqrt(sigma^2 + 1)  step :657-659 (v_prediction)
    sig2p1: float     # sigma^2 + 1
    dt: float         # sigma_next - sigma          step :670


class EulerSchedule:
    def __init__(self, num_inference_steps: int, sigma_min: float = 0.002, sigma_max: float = 700.0):
        if num_inference_steps < 1:
            raise ValueError("num_inference_steps must be >= 1")
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho = sigma_min ** (1 / rho)
        max_inv_rho = sigma_max ** (1 / rho)
        sig = (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho
        sig = np.concatenate([sig, [0]]).astype(np.float32)
        self.sigmas = torch.from_numpy(sig).to(torch.float32)                       # [N+1], last = 0
        self.timesteps = torch.Tensor([0.25 * s.log() for s in self.sigmas[:-1]])   # [N]
        self.num_inference_steps = num_inference_steps

    @property
    def init_noise_sigma(self) -> torch.Tensor:
        """'leading' spacing: sqrt(sigma_max^2 + 1) (:262-268)."""
        return (self.sigmas.max() ** 2 + 1) ** 0.5

    def index_for_timestep(self, t: torch.Tensor) -> int:
        """Position of t in timesteps; second match if duplicated (:554-566)."""
        idx = (self.timesteps == t).nonzero()
        return int(idx[1 if len(idx) > 1 else 0].item())

    def step_scalars(self, i: int) -> StepScalars:
        s = self.sigmas[i]
        return StepScalars(sigma=float(s), div=float((s ** 2 + 1) ** 0.5), c_out=float(-s / (s ** 2 + 1) ** 0.5),
                           sig2p1=float(s ** 2 + 1), dt=float(self.sigmas[i + 1] - s))

    def all_steps(self) -> List[StepScalars]:
        return [self.step_scalars(i) for i in range(self.num_inference_steps)]
