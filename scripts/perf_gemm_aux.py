"""Time the AlphaBlender layer class (linear + residual + aux operand, s_acc/s_res/s_aux) at the three UNet levels."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops
DEV = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
for (m, c) in [(258048, 320), (64512, 640), (16128, 1280)]:
    x = torch.randn(m, 4 * c, device=DEV).half()
    res = torch.randn(m, c, device=DEV).half()
    aux = torch.randn(m, c, device=DEV).half()
    gw = ops.prep_linear(torch.randn(c, 4 * c) * (4 * c) ** -0.5, torch.randn(c), DEV)
    out = torch.empty(m, c, device=DEV, dtype=torch.half)
    fn = lambda: ops.linear(x, gw, res=res, aux=aux, s_acc=0.6, s_res=0.6, s_aux=0.4, out=out)
    ts = []
    for i in range(7):
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print(f"linear+res+aux M={m} K={4 * c} N={c}: {ts[len(ts) // 2]:.4f} ms", flush=True)
