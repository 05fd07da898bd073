"""Time the spatial/temporal attention and norm kernels at north-star shapes, with a cuBLAS GEMM as clock reference."""
import os, sys, json
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
        flush.zero_(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


x = torch.randn(8192, 8192, device=DEV).half(); w = torch.randn(8192, 8192, device=DEV).half()
ref = timeit(lambda: torch.matmul(x, w.t()))
print(f"cublas 8192^3: {ref:.4f} ms ({2*8192**3/ref/1e9:.0f} TFLOP/s)", flush=True)
del x, w


def sdpa_rows():
    """The library bar on the same box: F.scaled_dot_product_attention (what the reference calls, attention_processor.py:2383)
    at the level-0 north-star shape, cuDNN and flash backends, plus flash_attn / flashinfer when they run on sm_100."""
    import torch.nn.functional as Fn
    from torch.nn.attention import SDPBackend, sdpa_kernel
    imgs, heads, s, d = 28, 5, 9216, 64
    q, k, v = (torch.randn(imgs, heads, s, d, device=DEV).half() for _ in range(3))
    fl = 4.0 * imgs * heads * s * s * d
    for name, be in (("cudnn", SDPBackend.CUDNN_ATTENTION), ("flash", SDPBackend.FLASH_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
        try:
            with sdpa_kernel(be):
                ms = timeit(lambda: Fn.scaled_dot_product_attention(q, k, v))
            print(f"torch SDPA [{name}] imgs={imgs} heads={heads} S={s} d={d}: {ms:.4f} ms  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
        except Exception as e:
            print(f"torch SDPA [{name}]: unavailable ({str(e).splitlines()[0][:120]})", flush=True)
    try:
        from flash_attn import flash_attn_func
        qf, kf, vf = (t.transpose(1, 2).contiguous() for t in (q, k, v))
        ms = timeit(lambda: flash_attn_func(qf, kf, vf))
        print(f"flash_attn {__import__('flash_attn').__version__} same shape: {ms:.4f} ms  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
    except Exception as e:
        print(f"flash_attn: unavailable ({str(e).splitlines()[0][:120]})", flush=True)


if ("sdpa" in sys.argv or "attn" in sys.argv) and "nosdpa" not in sys.argv:
    sdpa_rows()
NORM_ONLY = "norm" in sys.argv
ATTN_ONLY = "attn" in sys.argv
for (imgs, s, c) in [(28, 9216, 320), (28, 2304, 640), (28, 576, 1280), (28, 144, 1280), (28, 9216, 960)]:
    qkv = torch.randn(imgs * s, 3 * c, device=DEV).half()
    out = torch.empty(imgs * s, c, device=DEV, dtype=torch.half)
    if not NORM_ONLY and c <= 1280 and s >= 576:
        ms = timeit(lambda: ops.attention_spatial(qkv, imgs, s, c, out=out))
        fl = 4.0 * imgs * (c // 64) * s * s * 64
        print(f"attn_spatial imgs={imgs} S={s} C={c}: {ms:.4f} ms  {fl/ms/1e9:.0f} TFLOP/s", flush=True)
        ms = timeit(lambda: ops.attention_temporal(qkv, 2, 14, s, c, out=out))
        print(f"attn_temporal S={s} C={c}: {ms:.4f} ms  {(qkv.numel()+out.numel())*2/ms/1e6:.0f} GB/s", flush=True)
    if ATTN_ONLY:
        continue
    xin = torch.randn(imgs * s, c, device=DEV).half()
    nw = ops.prep_norm(torch.ones(c), torch.zeros(c), DEV)
    ms = timeit(lambda: ops.layer_norm(xin, nw, out=out))
    print(f"layernorm M={imgs*s} C={c}: {ms:.4f} ms  {2*xin.numel()*2/ms/1e6:.0f} GB/s", flush=True)
    ms = timeit(lambda: ops.group_norm(xin, nw, imgs, 1e-5, True, out=out))
    print(f"groupnorm spatial: {ms:.4f} ms  {3*xin.numel()*2/ms/1e6:.0f} GB/s (3 passes)", flush=True)
    ms = timeit(lambda: ops.group_norm(xin, nw, 2, 1e-5, True, out=out))
    print(f"groupnorm temporal: {ms:.4f} ms  {3*xin.numel()*2/ms/1e6:.0f} GB/s (3 passes)", flush=True)
    del qkv, out, xin
