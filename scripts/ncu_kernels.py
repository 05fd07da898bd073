"""Targeted launches for `ncu --set full`: a few representative shapes of the hot kernels (1 warm-up + 1 profiled each)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops
DEV = "cuda"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
m, c = 258048, 320
x = torch.randn(m, c, device=DEV).half()
if which in ("all", "gemm"):
    # order of conv_gemm_kernel launches (each twice): conv3x3 320->320 | qkv K=320 N=960 | geglu K=320 N=2560 | ff2+res K=1280 N=320
    xi = torch.randn(28, 72, 128, 320, device=DEV).half()
    gw = ops.prep_conv(torch.randn(320, 320, 3, 3) * 0.02, torch.randn(320), DEV)
    for _ in range(2):
        ops.conv3x3(xi, gw)
    gq = ops.prep_linear(torch.randn(960, c) * 0.05, None, DEV)
    for _ in range(2):
        ops.linear(x, gq)
    gg = ops.prep_linear(torch.randn(2560, c) * 0.05, torch.randn(2560), DEV, geglu=True)
    for _ in range(2):
        ops.linear(x, gg)
    x4 = torch.randn(m, 4 * c, device=DEV).half()
    g2 = ops.prep_linear(torch.randn(c, 4 * c) * 0.03, torch.randn(c), DEV)
    for _ in range(2):
        ops.linear(x4, g2, res=x)
if which in ("all", "attn"):
    qkv = torch.randn(4 * 9216, 3 * c, device=DEV).half()
    for _ in range(2):
        ops.attention_spatial(qkv, 4, 9216, c)
torch.cuda.synchronize()
print("done")
