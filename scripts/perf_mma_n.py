"""tcgen05.mma cost as a function of the N extent: long-K GEMM with a single N tile (bn = N), timed with
WIW_GEMM_DBG=1 (no TMA loads: MMA issue + epilogue only) and normally."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wiw_b200  # noqa
from wiw_b200 import ops
from wiw_b200 import lib as L
import ctypes

DEV = "cuda"
tag = f"dbg={os.environ.get('WIW_GEMM_DBG', '0')} cg={os.environ.get('WIW_GEMM_CG', 'auto')}"
m, k = 148 * 128 * 16, 2048
x = torch.randn(m, k, device=DEV).half()
for n in (32, 128, 160, 256, 320):
    gw = ops.prep_linear(torch.randn(n, k) * k ** -0.5, None, DEV)
    out = torch.empty(m, gw.n, device=DEV, dtype=torch.half)
    for _ in range(2):
        ops.linear(x, gw, out=out)
    ts = []
    for _ in range(5):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.linear(x, gw, out=out); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    mmas_per_sm = (m // 128 // 148) * (gw.n // gw.bn) * (k // 16)
    print(f"{tag} | N={n} bn={gw.bn} | {ms:.4f} ms | {2.0 * m * k * n / ms / 1e9:.1f} TFLOP/s | {ms * 1e6 / mmas_per_sm:.1f} ns per MMA", flush=True)
    if int(os.environ.get("WIW_GEMM_DBG", "0")) & 4:
        buf = (ctypes.c_longlong * 8)()
        L.require_device().wiw_debug_gemm_prof(buf)
        v = list(buf)
        steps = max(v[3], 1)
        print(f"     CTA0 cycles: mma warp total {v[0]} ({v[0] / steps:.0f}/k-step), waiting full {v[1]} ({v[1] / steps:.0f}/k-step), waiting acc-empty {v[2]};"
              f" producer total {v[4]}, waiting empty {v[5]}; epilogue warp total {v[7]}, waiting acc-full {v[6]}; {ms * 1e6 / max(v[0], 1):.3f} ns/cycle", flush=True)
