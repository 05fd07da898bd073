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
    # the reference's second quantisation, (clip(t) * 255).astype(uint8), TRUNCATES: it equals the requantisation LUT applied to
    # PIL's bytes (never above them, sometimes one below) -- what wiw_resize_bicubic_u8 reproduces on the GPU
    from wiw_b200 import resize as R
    q = (np.clip(out.numpy(), 0, 1) * 255).astype(np.uint8)
    for t in range(2):
        pil = np.asarray(Image.fromarray(fr[0, t].transpose(1, 2, 0)).resize((12, 10), Image.BICUBIC)).transpose(2, 0, 1)
        assert np.array_equal(q[0, t], R.requant_lut()[pil])
        d = pil.astype(int) - q[0, t].astype(int)
        assert d.min() >= 0 and d.max() <= 1


def test_latent_noise_matches_reference(golden_dir):
    """prepare_latents / sample_latent_noise / rotate_by_degrees / decode_action_seq_frames (pipeline :325-362, :750-844) against
    outputs of the reference's own functions (tests/golden/latent_noise.npz, oracle/make_golden_northstar_rollout.py): bit-exact."""
    from wiw_b200.pipeline import sample_latent_noise_
    from wiw_b200.scheduler import EulerSchedule
    g = np.load(os.path.join(golden_dir, "latent_noise.npz"))
    aids = W.action_ids_idx_encode(g["actions"])
    diag = torch.diagonal(aids, dim1=-2, dim2=-1)
    assert np.array_equal(diag.numpy(), g["decoded"].astype(np.int64))            # decode_action_seq_frames; frame 0 reads 'stop'
    for name, dt in (("f32", torch.float32), ("f16", torch.float16)):
        ref = g[f"noise_{name}"]
        raw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(int(g[f"seed_{name}"])), dtype=dt)
        for i in range(raw.shape[0]):
            sample_latent_noise_(raw[i], diag[i])
        assert np.array_equal(raw.float().numpy(), ref), name
    ref = g["prepared_f16"]
    raw = torch.randn(ref.shape, generator=torch.Generator().manual_seed(int(g["prepared_seed"])), dtype=torch.float16)
    for i in range(raw.shape[0]):
        sample_latent_noise_(raw[i], diag[i])
    lat = raw * EulerSchedule(25).init_noise_sigma                                # pipeline :361, what Navigator.inference does
    assert lat.dtype == torch.float16 and np.array_equal(lat.float().numpy(), ref)
    with pytest.raises(AssertionError):
        sample_latent_noise_(torch.zeros(3, 4, 2, 24), torch.tensor([4, 2, 3]))   # width % 16 != 0 (:817-818)


def test_error_policy(tmp_path):
    """Only per-request DATA errors are absorbed (explicit `errors` marker); configuration / library errors end the worker."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers"))
    import stub_wm_worker as S
    args = W.build_parser().parse_args(["--width", "64", "--height", "32", "--out_width", "24", "--out_height", "20", "--num_frames", "8",
                                        "--action_input_channel", "8"])
    worker = W.WorldModelWorker(args, S.FakeNavigator(args))
    worker._resizer_key = (args.height, args.width, args.out_height, args.out_width)
    worker._resizer = S.cpu_resizer(args)
    acts = np.array([[4, 1, 2, 2, 1, 3, 1, 1]], dtype=np.int64)
    ok = {"b_action": acts, "b_image": np.zeros((1, 3, 32, 64), np.uint8), "save_dirs": [str(tmp_path / "a")],
          "request_model_name": "igen", "return_objects": [True]}
    assert "errors" not in worker.do_some_tasks(ok)
    nofile = {k: v for k, v in ok.items() if k != "b_image"}
    bad = worker.do_some_tasks(nofile)                                             # no cond_rgb.png on disk
    W.check_outputdict(bad)
    assert bad["errors"] and len(bad["errors"]) == 1 and int(bad["pred_frames"].max()) == 0
    files_bad = worker.do_some_tasks({k: v for k, v in nofile.items() if k != "return_objects"})
    assert "pred_frames" not in files_bad and files_bad["errors"]                # files mode: no frames pretend to exist
    with pytest.raises(RuntimeError):                                              # third data error in a row: stop serving black frames
        worker.do_some_tasks(nofile)
    worker._consecutive_data_errors = 0
    with pytest.raises(ValueError):
        worker.do_some_tasks({**ok, "request_model_name": "cosmos"})
    with pytest.raises(RuntimeError):                                              # library-side failure (stub raises for action sum 77)
        worker.do_some_tasks({**ok, "b_action": np.array([[4, 73, 0, 0, 0, 0, 0, 0]], dtype=np.int64)})
    shape_bad = worker.do_some_tasks({**ok, "b_image": np.zeros((1, 2, 32, 64), np.uint8)})
    assert shape_bad["errors"]


def test_parser_accepts_reference_flags():
    a, unknown = W.build_parser().parse_known_args(
        "--width 1024 --height 576 --out_width 480 --out_height 480 --log_dir x --exp_id=e_igen --num_frames 14 --num_past_obs 1 "
        "--task_type navigation --action_strategy micro_cond --action_input_channel 14 --device cuda:0 --unet_path u --svd_path s "
        "--weight_dtype float16 --some_manager_flag=3".split())
    assert a.width == 1024 and a.exp_id == "e_igen" and unknown == ["--some_manager_flag=3"]
