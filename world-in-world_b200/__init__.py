"""wiw_b200 -- B200-native (sm_100a) implementation of the World-in-World SVD-dagger
world-model hot path: EDM/Euler denoising of a [B,F,C,H,W] latent video through the
spatio-temporal UNet + temporal-VAE decode, behind the reference's WM-server worker API.

The directory is named `world-in-world_b200` (not an identifier); import it as
`wiw_b200` through the shim module at the repo root.
"""
from .config import FULL_UNET, FULL_VAE, TINY_UNET, TINY_VAE, UNetConfig, VAEConfig  # noqa: F401

__all__ = ["UNetConfig", "VAEConfig", "FULL_UNET", "FULL_VAE", "TINY_UNET", "TINY_VAE"]
