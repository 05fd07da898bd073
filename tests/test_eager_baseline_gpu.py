"""The bar SURVEY.md 8(d) names beside the CPU baseline: the reference's graph in eager fp16 PyTorch on the same
B200.  The vendored fork cannot travel to the GPU box, so the oracle restatement of that graph (same ops, same
order: F.conv2d / F.linear / F.group_norm / F.scaled_dot_product_attention) is run on `cuda` in fp16 at the
north-star shape and timed with CUDA events next to this repo's engine in the same process.  The numbers land in
`gpurun_out/eager_fp16_unet.json`; the only assertions are finiteness, agreement of the two outputs and that the
engine is not slower than library PyTorch."""
import json
import os

import pytest
import torch

import wiw_b200  # noqa: F401
from oracle import svd_oracle as O
from wiw_b200 import weights as Wt
from wiw_b200.config import FULL_UNET
from wiw_b200 import ops
from wiw_b200.engine import UNetEngine, sinusoid

pytestmark = pytest.mark.gpu


def _time(fn, iters):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def test_unet_forward_vs_eager_fp16_pytorch():
    dev = torch.device("cuda")
    F_, h, w = 14, 72, 128
    sd = Wt.random_init_on_device(Wt.unet_param_spec(FULL_UNET), 0, dev)
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(2, F_, 8, h, w, generator=g, device=dev).half()
    enc = torch.randn(2, 1, 1024, generator=g, device=dev).half()
    enc[0] = 0
    ids = torch.tensor([[6.0, 127.0, 0.02]] * 2, device=dev).half()
    aids = O.action_ids_idx_encode(torch.tensor([[4] + [1, 2, 3] * 4 + [1]])).to(dev).half()
    cfg = O.UNetCfg()
    t = 1.25

    with torch.no_grad(), torch.device(dev):
        def eager():
            return O.unet_forward(sd, cfg, x, torch.tensor(t), enc, ids, aids)
        y_ref = eager()
        ms_eager = _time(eager, 3)
    assert torch.isfinite(y_ref).all()

    eng = UNetEngine(FULL_UNET, sd, dev)
    eng.set_condition(enc[1], 0.02, aids[0])
    xin = ops.nchw_to_nhwc(x.flatten(0, 1).contiguous(), 64)
    t_emb = sinusoid(torch.tensor([t]), FULL_UNET.block_out_channels[0]).half().to(dev)
    y = eng.forward(xin, t_emb)
    ms_wiw = _time(lambda: eng.forward(xin, t_emb), 5)
    y_nchw = ops.nhwc_to_nchw(y, 4).float().view(2, F_, 4, h, w)
    rel = ((y_nchw - y_ref.float()).pow(2).mean().sqrt() / y_ref.float().pow(2).mean().sqrt()).item()
    out = {"shape": [2, F_, 8, h, w], "eager_fp16_pytorch_ms": ms_eager, "wiw_b200_ms": ms_wiw,
           "speedup": ms_eager / ms_wiw, "rel_rms_vs_eager_fp16": rel, "torch": torch.__version__,
           "note": "eager = oracle restatement of the vendored UNet graph on cuda, fp16, library kernels (cuDNN/cuBLAS/SDPA)"}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/eager_fp16_unet.json", "w") as f:
        json.dump(out, f, indent=1)
    print(out)
    assert rel < 2e-2, rel
    assert ms_wiw < ms_eager
