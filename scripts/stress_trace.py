"""Find the GEMM launch whose output is not reproducible: run the full-size UNet forward (or VAE decode) several times with a checksum
after EVERY conv_gemm launch (this serialises the stream, so a race that needs back-to-back kernels may hide -- the summary says so)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops, weights as W
from wiw_b200.config import FULL_UNET, FULL_VAE
from wiw_b200.engine import UNetEngine, VAEDecoderEngine, sinusoid

DEV = "cuda"
WHAT = sys.argv[1] if len(sys.argv) > 1 else "unet"
RUNS = int(sys.argv[2]) if len(sys.argv) > 2 else 4
SYNC = "nosync" not in sys.argv
trace = []
orig = ops._launch


def traced(a, a_dim, a_stride, gw, out, out_dim, out_stride, taps, bx, by, res, aux, tab, tab_div, tab_mod, s_acc, s_res, s_aux, use_bias=True, act=0, bz=1):
    orig(a, a_dim, a_stride, gw, out, out_dim, out_stride, taps, bx, by, res, aux, tab, tab_div, tab_mod, s_acc, s_res, s_aux, use_bias, act, bz)
    rows = int(out_dim[1]) * int(out_dim[2]) * int(out_dim[3])
    desc = (f"M={rows} K={gw.k} N={gw.n} taps={len(taps)} bx={bx} by={by} bz={bz}{' geglu' if gw.geglu else ''}{' res' if res is not None else ''}"
            f"{' aux' if aux is not None else ''}{' tab' if tab is not None else ''}{' nobias' if (gw.bias is None or not use_bias) else ''} act={act} s_acc={s_acc:.3g}"
            f" out_stride={tuple(int(s) for s in out_stride)}")
    if SYNC:
        trace.append((desc, int(out.contiguous().view(torch.int16).to(torch.int64).sum().item())))
    else:
        trace.append((desc, out.contiguous().view(torch.int16).to(torch.int64).sum()))


ops._launch = traced
F, h, w = 14, 72, 128
if WHAT == "unet":
    eng = UNetEngine(FULL_UNET, W.random_init_on_device(W.unet_param_spec(FULL_UNET), 0, DEV), DEV)
    eng.set_condition(torch.randn(1, 1024), 0.02, torch.randint(0, 5, (F, 14)).float())
    x_in = torch.zeros(2 * F, h, w, 64, dtype=torch.float16, device=DEV)
    x_in[..., :8] = torch.randn(2 * F, h, w, 8, device=DEV).half()
    t_emb = sinusoid(torch.tensor([1.2]), 320).half().to(DEV)
    run = lambda: eng.forward(x_in, t_emb)
else:
    eng = VAEDecoderEngine(FULL_VAE, W.random_init_on_device(W.vae_decoder_param_spec(FULL_VAE), 1, DEV), DEV)
    z = torch.zeros(F, h, w, 64, dtype=torch.float16, device=DEV)
    z[..., :4] = torch.randn(F, h, w, 4, device=DEV).half()
    u8 = torch.empty(F, 3, 8 * h, 8 * w, dtype=torch.uint8, device=DEV)
    run = lambda: eng.decode(z, F, out_u8=u8)
runs = []
for r in range(RUNS):
    trace.clear()
    run()
    torch.cuda.synchronize()
    runs.append([(d, int(c)) for d, c in trace])
print(f"{WHAT}: {len(runs[0])} GEMM launches per run, per-launch sync {'ON' if SYNC else 'OFF'}")
bad = {}
for r in range(1, RUNS):
    for i, ((d0, c0), (d1, c1)) in enumerate(zip(runs[0], runs[r])):
        if c0 != c1:
            bad.setdefault(i, d0)
            break                      # only the FIRST divergence of a run is meaningful (later ones inherit it)
for i, d in sorted(bad.items()):
    print(f"first divergence at launch #{i}: {d}")
if not bad:
    print("all runs identical launch by launch")
