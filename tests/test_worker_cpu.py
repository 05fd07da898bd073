"""CPU tests of the WM-server boundary logic in wiw_b200/wm_worker.py: wire framing + worker_main loop with a fake
task_fn, request validation, action encoding against the golden fixture, output conversion semantics."""
import io
import os
import pickle
import threading

import numpy as np
import pytest
import torch

import wiw_b200  # noqa: F401
from wiw_b200 import wm_worker as W


def _frame(obj):
    d = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    return len(d).to_bytes(4, "big") + d


def test_framing_and_worker_main_loop():
    seen = []

    def task_fn(payload):
        seen.append(payload["x"])
        return {"pred_frames": np.full((1, 2, 3, 4, 4), payload["x"], dtype=np.uint8), "save_dirs": ["d"]}

    stdin = io.BytesIO(_frame((5, 0, {"x": 1})) + _frame((5, 1, {"x": 2})) + _frame("DONE"))
    r_fd, w_fd = os.pipe()
    t = threading.Thread(target=W.worker_main, args=(w_fd, task_fn, stdin))
    t.start()
    rf = os.fdopen(r_fd, "rb")
    a = W.read_frame(rf)
    b = W.read_frame(rf)
    t.join(20)
    assert not t.is_alive() and seen == [1, 2]                       # strictly serial, in order
    assert a[:2] == (5, 0) and b[:2] == (5, 1) and int(b[2]["pred_frames"][0, 0, 0, 0, 0]) == 2
    with pytest.raises(EOFError):
        W.read_frame(io.BytesIO(b"\x00\x00"))


def test_request_checks_and_name_matching():
    ok = {"b_action": np.zeros((1, 14), np.int64), "save_dirs": ["a"], "request_model_name": "igen",
          "b_image": np.zeros((1, 3, 8, 8), np.uint8), "return_objects": [True]}
    W.check_inputdict(ok)
    with pytest.raises(KeyError):
        W.check_inputdict({"b_action": [[1]], "save_dirs": ["a"]})
    with pytest.raises(AssertionError):
        W.check_inputdict({**ok, "b_image": np.zeros((1, 3, 8, 8), np.float32)})
    assert W.name_matches("igen", "igen") and W.name_matches("i", "igen") and W.name_matches("ig", "igen")
    assert W.name_matches("g", "igen") and W.name_matches("en", "igen") and W.name_matches("", "igen")     # Batcher slices
    assert not W.name_matches("cosmos", "igen") and not W.name_matches("c", "igen") and not W.name_matches(None, "igen")
    with pytest.raises(AssertionError):
        W.check_outputdict({"pred_frames": np.zeros((1,), np.float32), "save_dirs": []})
    with pytest.raises(AssertionError):
        W.check_outputdict({"video_tensors": 1, "save_dirs": []})


def test_action_ids_encode_matches_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "action_ids.npz"))
    enc = W.action_ids_idx_encode(g["actions"])
    assert enc.dtype == torch.int64 and (enc.numpy() == g["encoded"]).all()
    with pytest.raises(AssertionError):
        W.action_ids_idx_encode(np.zeros((2, 3, 4)))


def test_images_to_tensor_and_quantisation():
    """PIL BICUBIC resize then to_tensor (/255), then (clip*255).astype(uint8) truncation (api_models/__init__.py:159-166,229)."""
    from PIL import Image
    rng = np.random.default_rng(0)
    fr = rng.integers(0, 256, size=(1, 2, 3, 16, 32), dtype=np.uint8)
    out = W.images_to_tensor(fr, save_size=(12, 10))
    assert out.shape == (1, 2, 3, 10, 12) and out.dtype == torch.float32
    ref = np.asarray(Image.fromarray(fr[0, 1].transpose(1, 2, 0)).resize((12, 10), Image.BICUBIC), dtype=np.float32).transpose(2, 0, 1) / 255.0
    assert np.array_equal(out[0, 1].numpy(), ref)
    q = (np.clip(out.numpy(), 0, 1) * 255).astype(np.uint8)
    assert np.abs(q.astype(int) - np.asarray(Image.fromarray(fr[0, 0].transpose(1, 2, 0)).resize((12, 10), Image.BICUBIC)).transpose(2, 0, 1)[None, None][0, 0].astype(int)).max() <= 255


def test_parser_accepts_reference_flags():
    a, unknown = W.build_parser().parse_known_args(
        "--width 1024 --height 576 --out_width 480 --out_height 480 --log_dir x --exp_id=e_igen --num_frames 14 --num_past_obs 1 "
        "--task_type navigation --action_strategy micro_cond --action_input_channel 14 --device cuda:0 --unet_path u --svd_path s "
        "--weight_dtype float16 --some_manager_flag=3".split())
    assert a.width == 1024 and a.exp_id == "e_igen" and unknown == ["--some_manager_flag=3"]
