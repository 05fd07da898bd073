"""Cycle budget of one epilogue warp of conv_gemm_kernel (CTA 0, warp 3) per 32-column chunk, by phase (WIW_GEMM_DBG=4):
TMEM load + bias/table | activation math | accumulator release | residual | wait for the staging buffer | STS + proxy fence | TMA issue."""
import ctypes, os, sys
os.environ["WIW_GEMM_DBG"] = str(4 | int(os.environ.get("EXTRA_DBG", "0")))
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import lib as L, ops
DEV = "cuda"
lib = ctypes.CDLL(L.LIB_PATH)
names = ["bias add", "math", "release(+prefetch)", "residual", "wait_read", "sts+fence", "tma issue", "tmem ld+wait"]


def report(tag, fn, flop):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    buf = (ctypes.c_longlong * 24)()
    lib.wiw_debug_gemm_prof(buf)
    v = list(buf)
    ch = max(v[16], 1)
    print(f"{tag}: {ms:.4f} ms {flop / ms / 1e9:.0f} TFLOP/s | MMA warp: total {v[0]} cyc, wait data {v[1]}, wait acc-empty {v[2]} | epilogue warp: total {v[7]}, "
          f"wait acc-full {v[6]}, chunks {v[16]}")
    print("      cycles per chunk: " + ", ".join(f"{n} {v[8 + i] / ch:.0f}" for i, n in enumerate(names)) + f", first use of chunk a {v[17] / ch:.0f}, of chunk b {v[18] / ch:.0f} | sum {(sum(v[8:16]) + v[17] + v[18]) / ch:.0f}", flush=True)


m, c = 258048, 320
x = torch.randn(m, c, device=DEV).half()
res = torch.randn(m, c, device=DEV).half()
for (name, nn, geglu, r) in [("qkv K=320 N=960", 3 * c, False, None), ("geglu K=320 N=2560", 8 * c, True, None), ("proj+res K=320 N=320", c, False, res)]:
    gw = ops.prep_linear(torch.randn(nn, c) * c ** -0.5, torch.randn(nn), DEV, geglu=geglu)
    out = torch.empty(m, gw.n // 2 if geglu else gw.n, device=DEV, dtype=torch.half)
    report(name, lambda: ops.linear(x, gw, res=r, out=out), 2.0 * m * c * nn)
xi = torch.randn(28, 72, 128, 320, device=DEV).half()
gw = ops.prep_conv(torch.randn(320, 320, 3, 3) * 0.02, torch.randn(320), DEV)
report("conv3x3 320->320 level 0", lambda: ops.conv3x3(xi, gw), 2.0 * 258048 * 320 * 2880)
xv = torch.randn(2, 576, 1024, 128, device=DEV).half()
gv = ops.prep_conv(torch.randn(128, 128, 3, 3) * 0.03, torch.randn(128), DEV)
report("VAE conv3x3 128->128 2 frames 576x1024", lambda: ops.conv3x3(xv, gv), 2.0 * 2 * 576 * 1024 * 128 * 1152)
xt = torch.randn(1, 4, 576 * 1024, 128, device=DEV).half()
gt = ops.prep_conv(torch.randn(128, 128, 3, 1, 1) * 0.05, torch.randn(128), DEV)
report("VAE tconv3 128->128 4 frames 576x1024", lambda: ops.tconv3(xt, gt), 2.0 * 4 * 576 * 1024 * 128 * 384)
