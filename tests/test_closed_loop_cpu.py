"""BASELINE configs[3] host logic on the CPU: the closed-loop harness (wiw_b200/closed_loop.py: MiniManager + HeuristicSolver)
around the REAL worker process loop (wm_worker.worker_main + WorldModelWorker.do_some_tasks) with the CPU stub Navigator of
tests/helpers/stub_wm_worker.py.  Checks the wire protocol in both directions, the Batcher slicing / recomposition, that every
request depends on the previous answer, and that the loop is deterministic."""
import os
import sys

import numpy as np
import pytest

import wiw_b200  # noqa: F401
from wiw_b200 import closed_loop as CL

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
STUB = [sys.executable, os.path.join(HERE, "helpers", "stub_wm_worker.py"), "--width", "64", "--height", "32", "--out_width", "24",
        "--out_height", "20", "--num_frames", "8", "--action_input_channel", "8"]


def test_split_and_recompose_follow_the_reference_batcher():
    req = {"b_action": np.arange(6).reshape(3, 2), "save_dirs": ["a", "b", "c"], "request_model_name": "igen", "return_objects": [True] * 3}
    parts = CL.split_batch(req, 1)
    assert [p["request_model_name"] for p in parts] == ["i", "g", "e"] and [p["save_dirs"] for p in parts] == [["a"], ["b"], ["c"]]
    parts2 = CL.split_batch(req, 2)
    assert [p["request_model_name"] for p in parts2] == ["ig", "en"] and parts2[1]["b_action"].shape == (1, 2)
    out = CL.recompose([{"pred_frames": np.zeros((1, 2, 3)), "save_dirs": ["a"]}, {"pred_frames": np.ones((1, 2, 3)), "save_dirs": ["b"]}])
    assert out["save_dirs"] == ["a", "b"] and np.asarray(out["pred_frames"]).shape == (2, 2, 3)


@pytest.mark.parametrize("batch_size", [1, 2])
def test_closed_loop_four_step_horizon(tmp_path, batch_size):
    mgr = CL.MiniManager(STUB, batch_size=batch_size, cwd=REPO)
    try:
        runs = []
        for rep in range(2):
            solver = CL.HeuristicSolver(mgr.host, mgr.port, str(tmp_path / f"r{rep}"), num_frames=8)
            obs = np.random.default_rng(0).integers(0, 256, size=(3, 32, 64), dtype=np.uint8)
            stats = solver.run(obs, horizon=4, branches=2)
            solver.close()
            runs.append(stats)
        s = runs[0]
        assert s["requests"] == 4 and s["frames"] == 4 * 2 * 8 and s["final_obs_shape"] == [3, 20, 24]
        assert len(s["chosen"]) == 4 and all(c in (0, 1) for c in s["chosen"])
        assert runs[0]["chosen"] == runs[1]["chosen"] and runs[0]["scores"] == runs[1]["scores"]        # deterministic loop
        assert mgr.sub_batches == 2 * 4 * (2 // batch_size)
        assert s["requests_per_s"] > 0
    finally:
        assert mgr.stop(timeout=60) == 0


def test_solver_action_rows_match_the_reference_layout():
    s = CL.HeuristicSolver.__new__(CL.HeuristicSolver)
    s.F = 14
    rows = s.action_rows([["forward"], ["turn_left", "turn_left"]])
    assert rows.dtype == np.int64 and rows.shape == (2, 14)
    assert rows[0].tolist() == [4] + [1] * 13 and rows[1].tolist() == [4, 2, 2] + [1] * 11
