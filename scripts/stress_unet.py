"""Stress the hot path for rare races: N full-size UNet forwards (+ VAE decodes) on the SAME input must be bit-identical (the kernels
have no atomics and a fixed summation order), and none may trap.  Usage: stress_unet.py [n_unet] [n_vae]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import weights as W
from wiw_b200.config import FULL_UNET, FULL_VAE
from wiw_b200.engine import UNetEngine, VAEDecoderEngine, sinusoid

DEV = "cuda"
n_unet = int(sys.argv[1]) if len(sys.argv) > 1 else 200
n_vae = int(sys.argv[2]) if len(sys.argv) > 2 else 20
unet = UNetEngine(FULL_UNET, W.random_init_on_device(W.unet_param_spec(FULL_UNET), 0, DEV), DEV)
F, h, w = 14, 72, 128
unet.set_condition(torch.randn(1, 1024), 0.02, torch.randint(0, 5, (F, 14)).float())
x_in = torch.zeros(2 * F, h, w, 64, dtype=torch.float16, device=DEV)
x_in[..., :8] = torch.randn(2 * F, h, w, 8, device=DEV).half()
t_emb = sinusoid(torch.tensor([1.2]), 320).half().to(DEV)


def checksum(t):
    return int(t.contiguous().view(torch.int16).to(torch.int64).sum().item())


ref, bad, t0 = None, 0, time.time()
for it in range(n_unet):
    c = checksum(unet.forward(x_in, t_emb))
    if ref is None:
        ref = c
    elif c != ref:
        bad += 1
        print(f"unet forward {it}: checksum {c} != {ref}", flush=True)
    if it % 50 == 49:
        print(f"unet {it + 1} forwards, {bad} mismatches, {time.time() - t0:.0f}s", flush=True)
print(f"UNET: {n_unet} forwards, {bad} mismatches", flush=True)
vae = VAEDecoderEngine(FULL_VAE, W.random_init_on_device(W.vae_decoder_param_spec(FULL_VAE), 1, DEV), DEV)
z = torch.zeros(F, h, w, 64, dtype=torch.float16, device=DEV)
z[..., :4] = torch.randn(F, h, w, 4, device=DEV).half()
u8 = torch.empty(F, 3, 8 * h, 8 * w, dtype=torch.uint8, device=DEV)
ref, vbad = None, 0
for it in range(n_vae):
    vae.decode(z, F, out_u8=u8)
    c = int(u8.to(torch.int64).sum().item())
    if ref is None:
        ref = c
    elif c != ref:
        vbad += 1
        print(f"vae decode {it}: checksum {c} != {ref}", flush=True)
print(f"VAE: {n_vae} decodes, {vbad} mismatches", flush=True)
sys.exit(1 if (bad or vbad) else 0)
