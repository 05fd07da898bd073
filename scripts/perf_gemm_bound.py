"""Which resource bounds conv_gemm_kernel?  Run with WIW_GEMM_DBG=0|1|2 (0 = normal, 1 = no TMA loads, 2 = no MMA
issue; results are garbage for 1/2, only the timing matters) and WIW_GEMM_CG=1|2."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
tag = f"dbg={os.environ.get('WIW_GEMM_DBG', '0')} cg={os.environ.get('WIW_GEMM_CG', 'auto')}"


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rec(name, ms, flop):
    print(f"{tag} | {name} | {ms:.4f} ms | {flop / ms / 1e9:.1f} TFLOP/s", flush=True)


for (n, h, w, cin, cout) in [(28, 72, 128, 320, 320), (28, 36, 64, 640, 640), (28, 18, 32, 1280, 1280)]:
    x = torch.randn(n, h, w, cin, device=DEV).half()
    gw = ops.prep_conv(torch.randn(cout, cin, 3, 3) * (9 * cin) ** -0.5, torch.randn(cout), DEV)
    out = torch.empty(n, h, w, gw.n, device=DEV, dtype=torch.half)
    rec(f"conv3x3 {n}x{h}x{w} {cin}->{cout} bn={gw.bn}", timeit(lambda: ops.conv3x3(x, gw, out=out)), 2.0 * n * h * w * 9 * cin * cout)
    del x, out
m, c = 258048, 320
x = torch.randn(m, c, device=DEV).half()
for (name, nn, geglu) in [("qkv", 3 * c, False), ("ff1_geglu", 8 * c, True)]:
    gw = ops.prep_linear(torch.randn(nn, c) * c ** -0.5, torch.randn(nn), DEV, geglu=geglu)
    out = torch.empty(m, gw.n // 2 if geglu else gw.n, device=DEV, dtype=torch.half)
    rec(f"linear {name} M={m} K={c} N={nn} bn={gw.bn}", timeit(lambda: ops.linear(x, gw, out=out)), 2.0 * m * c * nn)
m = k = n = 8192
x = torch.randn(m, k, device=DEV).half()
gw = ops.prep_linear(torch.randn(n, k) * k ** -0.5, None, DEV)
out = torch.empty(m, n, device=DEV, dtype=torch.half)
rec(f"linear square 8192^3 bn={gw.bn}", timeit(lambda: ops.linear(x, gw, out=out)), 2.0 * m * k * n)
