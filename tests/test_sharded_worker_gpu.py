"""BASELINE configs[2] on real GPUs: ONE worker process owning 2 GPUs (`--gpus 2`), spawned exactly as the manager spawns a
worker; rank 0 starts its helper rank, a 3-branch request is sharded 2 + 1, one NCCL all-gather returns the frames."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _send(f, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    f.write(len(data).to_bytes(4, "big") + data)
    f.flush()


def _recv(f):
    n = int.from_bytes(f.read(4), "big")
    return pickle.loads(f.read(n))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_worker_two_gpus(tmp_path):
    r_fd, w_fd = os.pipe()
    cmd = [sys.executable, os.path.join(REPO, "world-in-world_b200", "wm_worker.py"), "--tiny", "--random_init", "--gpus", "2",
           "--width", "128", "--height", "64", "--out_width", "48", "--out_height", "40", "--num_frames", "8",
           "--action_input_channel", "8", "--num_inference_steps", "2", str(w_fd)]
    p = subprocess.Popen(cmd, stdin=subprocess.PIPE, pass_fds=[w_fd], cwd=REPO)
    os.close(w_fd)
    rf = os.fdopen(r_fd, "rb")
    try:
        rng = np.random.default_rng(0)
        b_image = rng.integers(0, 256, size=(3, 3, 64, 128), dtype=np.uint8)
        b_action = np.array([[4, 1, 2, 2, 1, 3, 1, 1], [4, 3, 3, 1, 1, 1, 2, 0], [4, 1, 1, 1, 1, 1, 1, 1]], dtype=np.int64)
        req = {"b_action": b_action, "b_image": b_image, "save_dirs": [str(tmp_path / f"b{i}") for i in range(3)],
               "request_model_name": "igen", "return_objects": [True] * 3}
        _send(p.stdin, (1, 0, req))
        _, _, out = _recv(rf)
        pf = out["pred_frames"]
        assert pf.shape == (3, 8, 3, 40, 48) and pf.dtype == np.uint8 and out["save_dirs"] == req["save_dirs"]
        assert all(pf[i].std() > 1.0 for i in range(3))                     # every branch came back from its rank
        assert not np.array_equal(pf[0], pf[1])
        _send(p.stdin, (1, 1, req))                                          # same request again: the per-(request, branch) seeds differ
        _, _, out2 = _recv(rf)
        assert out2["pred_frames"].shape == pf.shape
        _send(p.stdin, "DONE")
        p.stdin.close()
        assert p.wait(timeout=120) == 0
    finally:
        if p.poll() is None:
            p.kill()
