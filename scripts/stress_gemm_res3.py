"""Race-rate of the residual layer classes under the current library (WIW_B200_LIB selects a variant)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 500
m, c = 258048, 320
x4 = torch.randn(m, 4 * c, device=DEV).half()
gw3 = ops.prep_linear(torch.randn(c, 4 * c) * (4 * c) ** -0.5, torch.randn(c), DEV)
o3 = torch.empty(m, c, device=DEV, dtype=torch.half)
r = torch.randn(m, c, device=DEV).half()
xi = torch.randn(28, 72, 128, 320, device=DEV).half()
ri = torch.randn(28, 72, 128, 320, device=DEV).half()
gwc = ops.prep_conv(torch.randn(320, 320, 3, 3) * 0.02, torch.randn(320), DEV)
oi = torch.empty(28, 72, 128, 320, device=DEV, dtype=torch.half)
for name, fn in [("linear K=1280 N=320 + res", lambda: ops.linear(x4, gw3, res=r, out=o3)), ("conv3x3 320->320 + res", lambda: ops.conv3x3(xi, gwc, res=ri, out=oi))]:
    ref, bad = None, 0
    for i in range(REPS):
        y = fn()
        if ref is None:
            ref = y.clone()
        elif not torch.equal(y, ref):
            bad += 1
    print(f"{'RACE' if bad else 'ok  '} {name}: {bad}/{REPS - 1} differ", flush=True)
