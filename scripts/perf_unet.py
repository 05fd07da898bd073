"""Time one full-size UNet forward (CFG pair, F=14, 72x128 latents) and the VAE decode on the GPU."""
import json
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
t0 = time.time()
sd = W.random_init_on_device(W.unet_param_spec(FULL_UNET), 0, DEV)
unet = UNetEngine(FULL_UNET, sd, DEV)
del sd
torch.cuda.synchronize()
print(f"unet built in {time.time() - t0:.1f}s, mem {torch.cuda.memory_allocated() / 2**30:.2f} GiB", flush=True)
F, h, w = 14, 72, 128
unet.set_condition(torch.randn(1, 1024), 0.02, torch.randint(0, 5, (F, 14)).float())
x_in = torch.zeros(2 * F, h, w, 64, dtype=torch.float16, device=DEV)
x_in[..., :8] = torch.randn(2 * F, h, w, 8, device=DEV).half()
t_emb = sinusoid(torch.tensor([1.2]), 320).half().to(DEV)
res = {}
for it in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    v = unet.forward(x_in, t_emb)
    e1.record()
    torch.cuda.synchronize()
    print(f"unet forward {it}: {e0.elapsed_time(e1):.2f} ms  finite={bool(torch.isfinite(v).all())} absmean={float(v[..., :4].float().abs().mean()):.4f}", flush=True)
    res[f"unet_fwd_ms_{it}"] = e0.elapsed_time(e1)
print(f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)

if "--vae" in sys.argv:
    vsd = W.random_init_on_device(W.vae_decoder_param_spec(FULL_VAE), 1, DEV)
    vae = VAEDecoderEngine(FULL_VAE, vsd, DEV)
    z = torch.zeros(F, h, w, 64, dtype=torch.float16, device=DEV)
    z[..., :4] = torch.randn(F, h, w, 4, device=DEV).half()
    u8 = torch.empty(F, 3, 8 * h, 8 * w, dtype=torch.uint8, device=DEV)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        vae.decode(z, F, out_u8=u8)
        e1.record()
        torch.cuda.synchronize()
        print(f"vae decode {it}: {e0.elapsed_time(e1):.2f} ms  mean={float(u8.float().mean()):.2f}", flush=True)
        res[f"vae_ms_{it}"] = e0.elapsed_time(e1)
    print(f"peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/perf_unet.json", "w"), indent=1)
