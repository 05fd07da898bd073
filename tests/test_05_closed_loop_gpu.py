"""BASELINE configs[3] on the GPU box: solver-AR-shaped closed loop -- heuristic policy, WM server backed by the sm_100a kernels,
4-step horizon, 2 action branches per step, synthetic Habitat-shaped RGB observation (480x480 uint8), full 1.53 B-parameter
architecture at 576x1024 px / 14 frames / 25 Euler steps with seeded random-init weights.

The worker is spawned by the in-repo MiniManager exactly as the reference manager spawns it (`<python> wm_worker.py <flags> <w_fd>`,
stdin pipe; downstream/utils/worker_manager.py:303-346) and the solver speaks the reference's socket framing
(downstream/solver_base.py:645-673).  Reports requests/s and frames/s in gpurun_out/closed_loop.json."""
import json
import os
import sys

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import closed_loop as CL

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("batch_size", [1])
def test_closed_loop_four_steps_full_size(tmp_path, batch_size):
    tiny = os.environ.get("WIW_CLOSED_LOOP_TINY", "0") == "1"          # development aid: the same loop on the tiny geometry
    flags = (["--tiny", "--width", "128", "--height", "64", "--num_frames", "8", "--action_input_channel", "8", "--num_inference_steps", "4"]
             if tiny else ["--num_inference_steps", "25"])
    F = 8 if tiny else 14
    cmd = [sys.executable, os.path.join(REPO, "world-in-world_b200", "wm_worker.py"), "--random_init", "--out_width", "480",
           "--out_height", "480", "--exp_id=closed_loop"] + flags
    mgr = CL.MiniManager(cmd, batch_size=batch_size, cwd=REPO)
    try:
        solver = CL.HeuristicSolver(mgr.host, mgr.port, str(tmp_path), num_frames=F)
        obs = np.random.default_rng(11).integers(0, 256, size=(3, 480, 480), dtype=np.uint8)
        warm = solver.look_ahead(obs, [["forward"]], step=-1)          # model load + CUDA-graph capture + library autotune
        assert warm["pred_frames"].shape == (1, F, 3, 480, 480) and "errors" not in warm
        solver.requests = 0
        stats = solver.run(obs, horizon=4, branches=2)
        solver.close()
    finally:
        rc = mgr.stop(timeout=120)
    assert rc == 0
    assert stats["requests"] == 4 and stats["frames"] == 4 * 2 * F and stats["final_obs_shape"] == [3, 480, 480]
    assert all(np.isfinite(s) and s > 0 for row in stats["scores"] for s in row)      # real frames, not the zero-frame reply
    stats["config"] = ("tiny" if tiny else "full 1.53B UNet + temporal VAE, 576x1024, 14 frames, 25 steps, CFG") + f", manager batch_size {batch_size}"
    stats["gpu"] = torch.cuda.get_device_name(0)
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    with open(os.path.join(REPO, "gpurun_out", "closed_loop.json"), "w") as f:
        json.dump(stats, f, indent=1)
    print(f"closed loop: {stats['requests_per_s']:.3f} requests/s, {stats['frames_per_s']:.2f} frames/s, chosen {stats['chosen']}")
