"""Drop-in boundary test on the GPU: spawn world-in-world_b200/wm_worker.py exactly as the reference manager does
(`<python> <script> <flags...> <w_fd>` with stdin=PIPE, pass_fds=[w_fd]; downstream/utils/worker_manager.py:303-346),
send framed requests (4-byte big-endian length + pickle of (client_id, task_id, payload)) and read the framed answers."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import wm_worker

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _send(f, obj):
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    f.write(len(data).to_bytes(4, "big") + data)
    f.flush()


def _recv(f):
    n = int.from_bytes(f.read(4), "big")
    return pickle.loads(f.read(n))


def test_worker_process_protocol(tmp_path):
    r_fd, w_fd = os.pipe()
    cmd = [sys.executable, os.path.join(REPO, "world-in-world_b200", "wm_worker.py"), "--tiny", "--random_init",
           "--width", "128", "--height", "64", "--out_width", "48", "--out_height", "40", "--num_frames", "8",
           "--action_input_channel", "8", "--num_inference_steps", "2", "--exp_id=test", str(w_fd)]
    p = subprocess.Popen(cmd, stdin=subprocess.PIPE, pass_fds=[w_fd], cwd=REPO)
    os.close(w_fd)
    rf = os.fdopen(r_fd, "rb")
    try:
        rng = np.random.default_rng(0)
        b_image = rng.integers(0, 256, size=(2, 4, 64, 128), dtype=np.uint8)          # RGBA like Habitat sensors
        b_action = np.array([[4, 1, 2, 2, 1, 3, 1, 1], [4, 3, 3, 1, 1, 1, 2, 0]], dtype=np.int64)
        req = {"b_action": b_action, "b_image": b_image, "save_dirs": [str(tmp_path / "a"), str(tmp_path / "b")],
               "request_model_name": "igen", "return_objects": [True, True]}
        _send(p.stdin, (7, 0, req))
        cid, tid, out = _recv(rf)
        assert (cid, tid) == (7, 0)
        wm_worker.check_outputdict(out)
        pf = out["pred_frames"]
        assert pf.shape == (2, 8, 3, 40, 48) and pf.dtype == np.uint8 and out["save_dirs"] == req["save_dirs"]
        assert pf.std() > 1.0                                                        # not the zero-frame error answer
        # a Batcher-sliced request (batch_size 1 slices every value, including the name string: worker_manager.py:457-458)
        req1 = {"b_action": b_action[:1], "b_image": b_image[:1], "save_dirs": req["save_dirs"][:1],
                "request_model_name": "i", "return_objects": [True]}
        _send(p.stdin, (7, 1, req1))
        cid, tid, out1 = _recv(rf)
        assert tid == 1 and out1["pred_frames"].shape == (1, 8, 3, 40, 48)
        # files mode: <save_dir>/<idx>.jpg + action_seq.json (saver.py:221-274)
        req2 = {"b_action": b_action[:1], "b_image": b_image[:1, :3], "save_dirs": [str(tmp_path / "c")], "request_model_name": "igen"}
        _send(p.stdin, (7, 2, req2))
        _, _, out2 = _recv(rf)
        assert "pred_frames" not in out2 and sorted(os.listdir(tmp_path / "c")) == sorted([f"{i}.jpg" for i in range(8)] + ["action_seq.json"])
        # a request whose DATA cannot be served (no b_image and no cond_rgb.png on disk) is absorbed: well-formed zero frames
        # plus an explicit per-branch `errors` list; the worker keeps serving
        bad = {"b_action": b_action[:1], "save_dirs": [str(tmp_path / "missing")], "request_model_name": "igen", "return_objects": [True]}
        _send(p.stdin, (7, 3, bad))
        _, _, out3 = _recv(rf)
        assert out3["pred_frames"].shape == (1, 8, 3, 40, 48) and int(out3["pred_frames"].max()) == 0
        assert isinstance(out3["errors"], list) and len(out3["errors"]) == 1 and "cond_rgb" in out3["errors"][0]
        _send(p.stdin, (7, 4, req1))
        _, _, out4 = _recv(rf)
        assert out4["pred_frames"].std() > 1.0 and "errors" not in out4
        # an observation that is not 576x1024-shaped (the solvers send 384x384 / 480x480) goes through the LANCZOS branch
        req5 = {**req1, "b_image": rng.integers(0, 256, size=(1, 3, 48, 48), dtype=np.uint8)}
        _send(p.stdin, (7, 5, req5))
        _, _, out5 = _recv(rf)
        assert out5["pred_frames"].shape == (1, 8, 3, 40, 48) and out5["pred_frames"].std() > 1.0
        # a request for ANOTHER model is a configuration error: the worker dies loudly like the reference's (api_models/__init__.py:189)
        _send(p.stdin, (7, 6, {**req1, "request_model_name": "cosmos"}))
        p.stdin.close()
        assert p.wait(timeout=60) != 0
    finally:
        if p.poll() is None:
            p.kill()


def test_first_frame_ids_and_determinism():
    """Same request twice on a fresh Navigator with injected latents -> identical frames (no hidden state but the generator)."""
    args = wm_worker.build_parser().parse_args(["--tiny", "--random_init", "--width", "128", "--height", "64", "--num_frames", "8",
                                                "--action_input_channel", "8", "--num_inference_steps", "2"])
    w = wm_worker.build_worker(args)
    acts = wm_worker.action_ids_idx_encode(np.array([[1, 1, 2, 3, 1, 1, 1, 1]]))
    assert int(acts[0, 0, 0]) == 4 and int(acts[0, 3, 3]) == 3 and int(acts[0, 2, 5]) == 0
    img = np.random.default_rng(1).integers(0, 256, size=(1, 3, 64, 128), dtype=np.uint8)
    lat = torch.randn(1, 8, 4, 8, 16, generator=torch.Generator().manual_seed(3)).half()
    w.nav.generator.manual_seed(1)
    a = w.nav.inference(acts, img, latents=lat)
    w.nav.generator.manual_seed(1)
    b = w.nav.inference(acts, img, latents=lat)
    assert a.shape == (1, 8, 3, 64, 128) and np.array_equal(a, b)
