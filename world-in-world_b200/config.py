"""Model geometry of the SVD-dagger world model (the one hot path this package serves).

Mirrors the constructor arguments the reference worker uses:
UNetSpatioTemporalConditionModel (FTsvd/diffusers-private/diffusers/models/unets/
unet_spatio_temporal_condition.py:72-97, kwargs :168-172) and the temporal VAE decoder
(models/autoencoders/autoencoder_kl_temporal_decoder.py:29-36).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple


@dataclass(frozen=True)
class UNetConfig:
    in_channels: int = 8
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    num_attention_heads: Tuple[int, ...] = (5, 10, 20, 20)
    cross_attention_dim: int = 1024
    addition_time_embed_dim: int = 256
    projection_class_embeddings_input_dim: int = 768
    num_frames: int = 14
    action_input_channel: int = 14
    action_strategy: str = "micro_cond"
    task_type: str = "navigation"

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def validate(self) -> None:
        for c, h in zip(self.block_out_channels, self.num_attention_heads):
            if c % 64 or c // h != 64:
                raise ValueError(f"kernels require head_dim 64 and C % 64 == 0, got C={c} heads={h}")
        if self.action_strategy != "micro_cond":
            raise ValueError("only the 'micro_cond' action strategy (the deployed WM-server setting) is built")


@dataclass(frozen=True)
class VAEConfig:
    latent_channels: int = 4
    out_channels: int = 3
    block_out_channels: Tuple[int, ...] = (128, 256, 512, 512)
    layers_per_block: int = 2
    scaling_factor: float = 0.18215

    def validate(self) -> None:
        for c in self.block_out_channels:
            if c % 64:
                raise ValueError(f"kernels require C % 64 == 0, got {c}")


# The named configurations of BASELINE.json.
FULL_UNET = UNetConfig()
FULL_VAE = VAEConfig()
# Shrunken geometry for fast parity tests (same topology; head_dim stays 64).
TINY_UNET = UNetConfig(block_out_channels=(64, 64, 128, 128), num_attention_heads=(1, 1, 2, 2),
                       cross_attention_dim=32, num_frames=8, action_input_channel=8)
TINY_VAE = VAEConfig(block_out_channels=(64, 64, 128, 128))
