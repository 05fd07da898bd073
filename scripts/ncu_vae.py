"""Short single-GPU workload for ncu: full temporal-VAE decoder, 14 frames of 72x128 latents (1 warm-up + 1 listed decode)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import weights as W
from wiw_b200.config import FULL_VAE
from wiw_b200.engine import VAEDecoderEngine
from wiw_b200.pipeline import Rollout
DEV = "cuda"
vae = VAEDecoderEngine(FULL_VAE, W.random_init_on_device(W.vae_decoder_param_spec(FULL_VAE), 1, DEV), DEV)
ro = Rollout.__new__(Rollout)
ro.vae, ro.device = vae, torch.device(DEV)
lat = torch.randn(14, 72, 128, 4, device=DEV).half() * 0.5
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for i in range(n):
    if i == n - 1:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    Rollout.decode(ro, lat)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
