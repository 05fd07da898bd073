"""Generate tests/golden/unet_northstar.npz: ONE UNet forward of the reference's own module at the BASELINE configs[1] shape.

    UNetSpatioTemporalConditionModel (full architecture, 1.53 B parameters) on sample [2, 14, 8, 72, 128] -- the CFG pair of a
    576x1024, 14-frame clip -- fp32 on the CPU (89.6 TFLOP: several minutes here).

Runs only in the build container (needs /root/reference).  Inputs and weights are seeded with numpy PCG64 so the GPU box
regenerates them bit-identically; the fixture holds only the seeds, the timestep and the output (1 MB).  Test infrastructure.

    python oracle/make_golden_northstar.py
"""
from __future__ import annotations

import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "oracle"))
import make_golden as MG  # noqa: E402  (sys.path for the vendored fork + the transformers-5 shim)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from wiw_b200 import weights as W  # noqa: E402
from wiw_b200.config import FULL_UNET  # noqa: E402

SEED_UNET, SEED_IN = 31, 32
F_, h, w = 14, 72, 128


def inputs():
    """Seeded inputs shared with tests/test_zz_northstar_forward_gpu.py."""
    sample = MG.rng_normal(SEED_IN, (2, F_, 8, h, w))
    enc = MG.rng_normal(SEED_IN + 1, (2, 1, FULL_UNET.cross_attention_dim))
    enc[0] = 0
    acts = torch.tensor([[4, 1, 2, 2, 1, 3, 1, 1, 2, 1, 1, 3, 3, 1]], dtype=torch.float32)
    return sample, enc, acts, torch.tensor(1.2345)


@torch.no_grad()
def main() -> None:
    torch.set_num_threads(os.cpu_count() or 1)
    c = FULL_UNET
    t0 = time.time()
    unet = MG.UNetSpatioTemporalConditionModel(
        block_out_channels=c.block_out_channels, num_attention_heads=c.num_attention_heads,
        cross_attention_dim=c.cross_attention_dim, action_strategy="micro_cond", task_type="navigation",
        num_frames=c.num_frames, action_input_channel=c.action_input_channel)
    unet.load_state_dict(W.random_init(W.unet_param_spec(c), SEED_UNET), strict=True)
    unet.eval()
    print(f"module built in {time.time() - t0:.0f} s", flush=True)
    sample, enc, acts, tstep = inputs()
    aids = MG.action_ids_idx_encode(1, acts).float()
    tids = torch.tensor([[6.0, 127.0, 0.02]] * 2)
    t1 = time.time()
    out = unet(sample, tstep, encoder_hidden_states=enc, added_time_ids=tids, return_dict=False, added_action_ids=aids)[0]
    print(f"forward ran in {time.time() - t1:.0f} s; out {tuple(out.shape)} std {float(out.std()):.4f} absmax {float(out.abs().max()):.3f}", flush=True)
    np.savez_compressed(os.path.join(MG.OUT, "unet_northstar.npz"), seed_unet=np.int64(SEED_UNET), seed_in=np.int64(SEED_IN),
                        actions=acts.numpy().astype(np.int64), timestep=np.float32(tstep.item()), out=out.numpy().astype(np.float16))      # fp16 storage: 2 MB, 3e-4 relative resolution


if __name__ == "__main__":
    main()
