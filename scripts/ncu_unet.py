"""Short single-GPU workload for ncu: build the full UNet, run 2 forwards (1 warm-up + 1 profiled window)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import weights as W
from wiw_b200.config import FULL_UNET
from wiw_b200.engine import UNetEngine, sinusoid
DEV = "cuda"
unet = UNetEngine(FULL_UNET, W.random_init_on_device(W.unet_param_spec(FULL_UNET), 0, DEV), DEV)
F, h, w = 14, 72, 128
unet.set_condition(torch.randn(1, 1024), 0.02, torch.randint(0, 5, (F, 14)).float())
x_in = torch.zeros(2 * F, h, w, 64, dtype=torch.float16, device=DEV)
x_in[..., :8] = torch.randn(2 * F, h, w, 8, device=DEV).half()
t_emb = sinusoid(torch.tensor([1.2]), 320).half().to(DEV)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
for _ in range(n):
    unet.forward(x_in, t_emb)
torch.cuda.synchronize()
print("done")
