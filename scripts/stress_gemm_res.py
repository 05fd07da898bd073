"""Many repetitions of the residual layer classes (and no-residual controls): any output not bit-identical to the first is a race."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 300
total = 0


def check(name, fn):
    global total
    ref, bad, first = None, 0, None
    for i in range(REPS):
        y = fn()
        if ref is None:
            ref = y.clone()
            continue
        if not torch.equal(y, ref):
            bad += 1
            if first is None:
                d = (y.float() - ref.float()).abs().view(-1, y.shape[-1])
                idx = torch.nonzero(d > 0)
                rows, cols = idx[:, 0], idx[:, 1]
                first = (f"{idx.shape[0]} elements; rows {int(rows.min())}..{int(rows.max())} ({len(set((rows // 128).tolist()))} M-tiles, rows%128 in "
                         f"{int((rows % 128).min())}..{int((rows % 128).max())}); cols {int(cols.min())}..{int(cols.max())} (32-col chunks {sorted(set((cols // 32).tolist()))}); max |diff| {float(d.max()):.4g}")
    total += bad
    print(f"{'RACE' if bad else 'ok  '} {name}: {bad}/{REPS - 1} differ" + (f" | first: {first}" if first else ""), flush=True)


m, c = 64512, 640
x = torch.randn(m, c, device=DEV).half()
r = torch.randn(m, c, device=DEV).half()
gw = ops.prep_linear(torch.randn(c, c) * c ** -0.5, torch.randn(c), DEV)
o = torch.empty(m, c, device=DEV, dtype=torch.half)
check("linear K=640 N=640 + res", lambda: ops.linear(x, gw, res=r, out=o))
check("linear K=640 N=640 (control, no res)", lambda: ops.linear(x, gw, out=o))
m, c = 258048, 320
x4 = torch.randn(m, 4 * c, device=DEV).half()
r3 = torch.randn(m, c, device=DEV).half()
gw3 = ops.prep_linear(torch.randn(c, 4 * c) * (4 * c) ** -0.5, torch.randn(c), DEV)
o3 = torch.empty(m, c, device=DEV, dtype=torch.half)
check("linear K=1280 N=320 + res", lambda: ops.linear(x4, gw3, res=r3, out=o3))
xi = torch.randn(28, 72, 128, 320, device=DEV).half()
ri = torch.randn(28, 72, 128, 320, device=DEV).half()
gwc = ops.prep_conv(torch.randn(320, 320, 3, 3) * 0.02, torch.randn(320), DEV)
oi = torch.empty(28, 72, 128, 320, device=DEV, dtype=torch.half)
check("conv3x3 320->320 + res", lambda: ops.conv3x3(xi, gwc, res=ri, out=oi))
check("conv3x3 320->320 (control)", lambda: ops.conv3x3(xi, gwc, out=oi))
xv = torch.randn(2, 144, 256, 512, device=DEV).half()
rv = torch.randn(2, 144, 256, 512, device=DEV).half()
gv = ops.prep_conv(torch.randn(512, 512, 3, 3) * 0.01, torch.randn(512), DEV)
ov = torch.empty(2, 144, 256, 512, device=DEV, dtype=torch.half)
check("VAE conv3x3 512->512 + res", lambda: ops.conv3x3(xv, gv, res=rv, out=ov))
print(f"TOTAL {total}")
sys.exit(1 if total else 0)
