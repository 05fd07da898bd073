"""Parity at BASELINE's FULL size: one UNet forward of the 1.53 B-parameter architecture on the CFG pair of a 576x1024, 14-frame
clip (sample [2, 14, 8, 72, 128]) on the sm_100a kernels against the output of the reference's own
UNetSpatioTemporalConditionModel run in fp32 on the CPU (tests/golden/unet_northstar.npz, oracle/make_golden_northstar.py;
inputs and weights regenerated here from the same PCG64 seeds).

Tolerance: fp16 storage / fp32 accumulation over reductions up to K = 23 040 against an fp32 reference: relative RMS <= 2e-2
(the tiny-geometry forward measures 1.7e-3; the margin is recorded in gpurun_out/parity.jsonl).  Named zz_ so it runs last."""
import math
import os

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import ops
from wiw_b200 import weights as W
from wiw_b200.config import FULL_UNET
from wiw_b200.engine import UNetEngine, sinusoid
from wiw_b200.wm_worker import action_ids_idx_encode

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rng_normal(seed: int, shape) -> torch.Tensor:
    return torch.from_numpy(np.random.Generator(np.random.PCG64(seed)).standard_normal(shape, dtype=np.float32))


def test_full_size_unet_forward_vs_reference_module(golden_dir):
    path = os.path.join(golden_dir, "unet_northstar.npz")
    if not os.path.exists(path):
        pytest.skip("north-star golden not generated")
    g = np.load(path)
    F, h, w = 14, 72, 128
    seed_in = int(g["seed_in"])
    sample = rng_normal(seed_in, (2, F, 8, h, w))
    enc = rng_normal(seed_in + 1, (2, 1, FULL_UNET.cross_attention_dim))
    aids = action_ids_idx_encode(g["actions"])[0].float()
    unet = UNetEngine(FULL_UNET, W.random_init(W.unet_param_spec(FULL_UNET), int(g["seed_unet"])), DEV)
    unet.set_condition(enc[1], 0.02, aids)
    x_in = ops.nchw_to_nhwc(sample.flatten(0, 1).half().to(DEV).contiguous(), 64)
    t_emb = sinusoid(torch.tensor([float(g["timestep"])]), FULL_UNET.block_out_channels[0]).half().to(DEV)
    v = unet.forward(x_in, t_emb)
    torch.cuda.synchronize()
    out = ops.nhwc_to_nchw(v, 4).float().cpu().view(2, F, 4, h, w)
    ref = torch.from_numpy(g["out"].astype(np.float32))              # stored as fp16 (2 MB)
    assert torch.isfinite(out).all()
    rel = float(((out - ref).pow(2).mean() / ref.pow(2).mean()).sqrt())
    rng = float(ref.max() - ref.min())
    psnr = 20 * math.log10(rng / float((out - ref).pow(2).mean().sqrt()))
    from conftest import record_parity
    record_parity("northstar_unet_forward_vs_reference_module", rel_rms=rel, psnr_db=psnr)
    print(f"north-star UNet forward: rel rms {rel:.4g}, PSNR {psnr:.1f} dB")
    assert rel < 2e-2, f"rel rms {rel:.4g}"
    assert psnr > 45.0
