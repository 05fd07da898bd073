"""Where do the rare wrong values of residual layers come from?  (A) constant residual: a hazard on the residual slots becomes invisible.
(B) zero activations + zero bias + coded residual r[row, col] = (row % 32) * 64 + (col % 64): the output IS the residual, so a wrong
element names the (row % 32, col % 64) it was taken from."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 400
m, c = 258048, 320
x4 = torch.randn(m, 4 * c, device=DEV).half()
gw3 = ops.prep_linear(torch.randn(c, 4 * c) * (4 * c) ** -0.5, torch.randn(c), DEV)
o3 = torch.empty(m, c, device=DEV, dtype=torch.half)


def run(name, fn, decode=False):
    ref, bad = None, 0
    for i in range(REPS):
        y = fn()
        if ref is None:
            ref = y.clone()
            continue
        if not torch.equal(y, ref):
            bad += 1
            if bad <= 4:
                d = (y.float() - ref.float()).abs()
                idx = torch.nonzero(d > 0)
                rows, cols = idx[:, 0], idx[:, 1]
                msg = f"   rep {i}: {idx.shape[0]} elements, rows {int(rows.min())}..{int(rows.max())} (row%128 {int((rows % 128).min())}..{int((rows % 128).max())}), cols {int(cols.min())}..{int(cols.max())}"
                if decode:
                    r0, c0 = int(rows[0]), int(cols[0])
                    got, exp = float(y[r0, c0]), float(ref[r0, c0])
                    msg += f" | element ({r0}, {c0}): got {got:.0f} = (row%32 {int(got) // 64}, col%64 {int(got) % 64}), expected {exp:.0f} = (row%32 {r0 % 32}, col%64 {c0 % 64})"
                    r1, c1 = int(rows[-1]), int(cols[-1])
                    got, exp = float(y[r1, c1]), float(ref[r1, c1])
                    msg += f" | last ({r1}, {c1}): got {got:.0f} = ({int(got) // 64}, {int(got) % 64}), expected ({r1 % 32}, {c1 % 64})"
                print(msg, flush=True)
    print(f"{'RACE' if bad else 'ok  '} {name}: {bad}/{REPS - 1} differ", flush=True)


r_rand = torch.randn(m, c, device=DEV).half()
run("random residual (baseline)", lambda: ops.linear(x4, gw3, res=r_rand, out=o3))
r_const = torch.full((m, c), 0.5, device=DEV, dtype=torch.half)
run("(A) constant residual", lambda: ops.linear(x4, gw3, res=r_const, out=o3))
rows = torch.arange(m, device=DEV).view(-1, 1)
cols = torch.arange(c, device=DEV).view(1, -1)
r_code = ((rows % 32) * 64 + (cols % 64)).half()
xz = torch.zeros(m, 4 * c, device=DEV, dtype=torch.half)
gwz = ops.prep_linear(torch.randn(c, 4 * c) * 0.01, torch.zeros(c), DEV)
run("(B) zero activations, coded residual", lambda: ops.linear(xz, gwz, res=r_code, out=o3), decode=True)
xr = torch.randn(m, 4 * c, device=DEV).half()
run("(C) random activations, zero residual", lambda: ops.linear(xr, gw3, res=torch.zeros_like(r_rand), out=o3))
