"""Time the implicit-GEMM kernel on representative north-star shapes (CUDA events, L2 flushed)."""
import json
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops

DEV = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)


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


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""
rows = []
def rec(name, ms, flop):
    rows.append({"name": name, "ms": round(ms, 4), "tflops": round(flop / ms / 1e9, 1)})
    print(rows[-1], flush=True)

for (m, c) in [(258048, 320), (64512, 640), (16128, 1280)]:
    x = torch.randn(m, c, device=DEV).half()
    for (name, n, geglu) in [("qkv", 3 * c, False), ("proj", c, False), ("ff1_geglu", 8 * c, True)]:
        if ONLY and ONLY not in name:
            continue
        if ONLY == "res":
            continue
        gw = ops.prep_linear(torch.randn(n, c) * c ** -0.5, torch.randn(n), DEV, geglu=geglu)
        out = torch.empty(m, gw.n // 2 if geglu else gw.n, device=DEV, dtype=torch.half)
        rec(f"linear {name} M={m} K={c} N={n} bn={gw.bn}", timeit(lambda: ops.linear(x, gw, out=out)), 2.0 * m * c * n)
    if ONLY and ONLY != "res":
        continue
    x4 = torch.randn(m, 4 * c, device=DEV).half()
    res = torch.randn(m, c, device=DEV).half()
    gw = ops.prep_linear(torch.randn(c, 4 * c) * (4 * c) ** -0.5, torch.randn(c), DEV)
    out = torch.empty(m, c, device=DEV, dtype=torch.half)
    rec(f"linear ff2+res M={m} K={4*c} N={c} bn={gw.bn}", timeit(lambda: ops.linear(x4, gw, res=res, out=out)), 2.0 * m * 4 * c * c)
    gwp = ops.prep_linear(torch.randn(c, c) * c ** -0.5, torch.randn(c), DEV)
    rec(f"linear proj+res M={m} K={c} N={c} bn={gwp.bn}", timeit(lambda: ops.linear(x, gwp, res=res, out=out)), 2.0 * m * c * c)
    del x4, res

if ONLY:
    sys.exit(0)
for (n, h, w, cin, cout) in [(28, 72, 128, 320, 320), (28, 36, 64, 640, 640), (28, 18, 32, 1280, 1280), (28, 72, 128, 960, 320),
                             (28, 9, 16, 1280, 1280), (28, 9, 16, 2560, 1280), (28, 18, 32, 2560, 1280)]:
    x = torch.randn(n, h, w, cin, device=DEV).half()
    gw = ops.prep_conv(torch.randn(cout, cin, 3, 3) * (9 * cin) ** -0.5, torch.randn(cout), DEV)
    out = torch.empty(n, h, w, gw.n, device=DEV, dtype=torch.half)
    rec(f"conv3x3 {n}x{h}x{w} {cin}->{cout}", timeit(lambda: ops.conv3x3(x, gw, out=out)), 2.0 * n * h * w * 9 * cin * cout)
    del x, out

m = k = n = 8192
x = torch.randn(m, k, device=DEV).half()
wt = torch.randn(n, k) * k ** -0.5
gw = ops.prep_linear(wt, None, DEV)
out = torch.empty(m, n, device=DEV, dtype=torch.half)
rec(f"linear square 8192^3 bn={gw.bn}", timeit(lambda: ops.linear(x, gw, out=out)), 2.0 * m * k * n)
wd = wt.half().to(DEV)
rec("cublas 8192^3 (reference point)", timeit(lambda: torch.matmul(x, wd.t())), 2.0 * m * k * n)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/perf_gemm.json", "w"), indent=1)
