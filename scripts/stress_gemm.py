"""Per-shape determinism check of conv_gemm_kernel: each layer class is launched REPS times on the same input (L2 flushed in between to
vary timing); any output that is not bit-identical to the first is a data race.  Usage: stress_gemm.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 20
flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=DEV)
total_bad = 0


def check(name, fn):
    global total_bad
    ref, bad, worst = None, 0, 0.0
    for i in range(REPS):
        if i % 2:
            flush.zero_()
        y = fn()
        torch.cuda.synchronize()
        if ref is None:
            ref = y.clone()
        elif not torch.equal(y, ref):
            bad += 1
            d = (y.float() - ref.float()).abs()
            worst = max(worst, float(d.max()))
            if bad == 1:
                idx = torch.nonzero(d.view(-1, d.shape[-1]) > 0)
                rows, cols = idx[:, 0], idx[:, 1]
                print(f"   first mismatch: {idx.shape[0]} elements, rows {int(rows.min())}..{int(rows.max())} (mod 128: {sorted(set((rows % 128).tolist()))[:8]}...), "
                      f"cols {int(cols.min())}..{int(cols.max())}, distinct col%32 {len(set((cols % 32).tolist()))}, distinct col//32 {sorted(set((cols // 32).tolist()))[:12]}", flush=True)
    total_bad += bad
    print(f"{'RACE' if bad else 'ok  '} {name}: {bad}/{REPS - 1} differ, max abs diff {worst:.4g}", flush=True)


m, c = 258048, 320
x = torch.randn(m, c, device=DEV).half()
res = torch.randn(m, c, device=DEV).half()
tabv = torch.randn(28, c, device=DEV).half()
for (name, n, geglu) in [("linear qkv K=320 N=960", 3 * c, False), ("linear proj K=320 N=320", c, False), ("linear geglu K=320 N=2560", 8 * c, True)]:
    gw = ops.prep_linear(torch.randn(n, c) * c ** -0.5, torch.randn(n), DEV, geglu=geglu)
    out = torch.empty(m, gw.n // 2 if geglu else gw.n, device=DEV, dtype=torch.half)
    check(name, lambda: ops.linear(x, gw, out=out))
gwp = ops.prep_linear(torch.randn(c, c) * c ** -0.5, torch.randn(c), DEV)
outp = torch.empty(m, c, device=DEV, dtype=torch.half)
check("linear proj+res K=320 N=320", lambda: ops.linear(x, gwp, res=res, out=outp))
check("linear proj+res+tab K=320 N=320", lambda: ops.linear(x, gwp, res=res, tab=tabv, tab_div=9216, tab_mod=28, out=outp))
gwnb = ops.prep_linear(torch.randn(c, c) * c ** -0.5, None, DEV)
check("linear no-bias K=320 N=320", lambda: ops.linear(x, gwnb, out=outp))
x6 = torch.randn(64512, 640, device=DEV).half()
gw6 = ops.prep_linear(torch.randn(640, 640) * 640 ** -0.5, torch.randn(640), DEV)
o6 = torch.empty(64512, 640, device=DEV, dtype=torch.half)
r6 = torch.randn(64512, 640, device=DEV).half()
check("linear proj+res K=640 N=640 (bn=224, 7 chunks)", lambda: ops.linear(x6, gw6, res=r6, out=o6))
xs = torch.randn(4032, 1280, device=DEV).half()
gws = ops.prep_linear(torch.randn(1280, 1280) * 1280 ** -0.5, torch.randn(1280), DEV)
check("linear K=1280 N=1280 M=4032 (single-CTA tiles)", lambda: ops.linear(xs, gws))
xi = torch.randn(28, 72, 128, 320, device=DEV).half()
gwc = ops.prep_conv(torch.randn(320, 320, 3, 3) * 0.02, torch.randn(320), DEV)
check("conv3x3 320->320 level 0", lambda: ops.conv3x3(xi, gwc))
check("conv3x3 320->320 level 0 + tab", lambda: ops.conv3x3(xi, gwc, tab=tabv, tab_div=9216, tab_mod=28))
xv = torch.randn(2, 576, 1024, 128, device=DEV).half()
rv = torch.randn(2, 576, 1024, 128, device=DEV).half()
gv = ops.prep_conv(torch.randn(128, 128, 3, 3) * 0.03, torch.randn(128), DEV)
check("VAE conv3x3 128->128 2x576x1024", lambda: ops.conv3x3(xv, gv))
check("VAE conv3x3 128->128 + res", lambda: ops.conv3x3(xv, gv, res=rv))
xt = torch.randn(1, 4, 576 * 1024, 128, device=DEV).half()
rt = torch.randn(1, 4, 576 * 1024, 128, device=DEV).half()
gt = ops.prep_conv(torch.randn(128, 128, 3, 1, 1) * 0.05, torch.randn(128), DEV)
check("VAE tconv3 128->128 4 frames", lambda: ops.tconv3(xt, gt))
check("VAE tconv3 128->128 + res, s_acc=0.7", lambda: ops.tconv3(xt, gt, res=rt, s_acc=0.7, s_res=1.0))
xv5 = torch.randn(2, 144, 256, 512, device=DEV).half()
gv5 = ops.prep_conv(torch.randn(512, 512, 3, 3) * 0.01, torch.randn(512), DEV)
check("VAE conv3x3 512->512 2x144x256", lambda: ops.conv3x3(xv5, gv5))
gv3 = ops.prep_conv(torch.randn(3, 128, 3, 3) * 0.03, torch.randn(3), DEV)
check("VAE conv_out 128->3 (n padded to 32)", lambda: ops.conv3x3(xv, gv3))
print(f"TOTAL mismatching launches: {total_bad}")
sys.exit(1 if total_bad else 0)
