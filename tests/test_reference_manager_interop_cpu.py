"""Drop-in evidence on the CPU: the reference's OWN manager stack (downstream/utils/worker_manager.py -- create_workers,
receiver_thread, ClientHandler + Batcher, write_framed / read_framed) drives this repo's worker process loop and request
handling (wm_worker.worker_main + WorldModelWorker.do_some_tasks; only the Navigator is a CPU fake, tests/helpers/).

Runs where /root/reference exists (the build container); skipped on the GPU box."""
import os
import socket
import sys
import threading

import numpy as np
import pytest

REF = "/root/reference"
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "downstream", "utils")), reason="reference checkout not present")


@pytest.fixture(scope="module")
def RM():
    sys.path.insert(0, REF)
    try:
        from downstream.utils import worker_manager
    finally:
        sys.path.remove(REF)
    return worker_manager


def _expected(actions_row, t):
    a = np.array(actions_row).copy()
    a[0] = 4                                             # action_ids_idx_encode forces 'stop' in column 0
    return (int(a.sum()) * 7 + t) % 256


# batch_size -1 (no Batcher) cannot be exercised: the shipped ClientHandler._send_ready_results touches self.batcher
# unconditionally (worker_manager.py:614) and closes the connection on the first empty poll.
@pytest.mark.parametrize("batch_size", [1, 2])
def test_reference_manager_drives_our_worker(RM, tmp_path, batch_size):
    cmd = [sys.executable, os.path.join(REPO, "tests", "helpers", "stub_wm_worker.py"), "--width 64 --height 32 --out_width 24",
           "--out_height 20 --num_frames 8 --action_input_channel 8 --log_dir x --exp_id=e_igen --some_manager_flag=1"]
    workers = RM.create_workers(cmd, 2)
    manager = RM.ManagerState(workers)
    rx = threading.Thread(target=RM.receiver_thread, args=(manager.client_out_queues, workers, "world_model"), daemon=True)
    rx.start()
    a, b = socket.socketpair()
    handler = RM.ClientHandler(a, "test", 0, manager, batch_size, "world_model")
    handler.start()
    try:
        acts = np.array([[4, 1, 2, 2, 1, 3, 1, 1], [1, 3, 3, 1, 1, 1, 2, 0], [2, 2, 2, 2, 1, 1, 1, 1]], dtype=np.int64)
        req = {"b_action": acts, "b_image": np.zeros((3, 4, 32, 64), np.uint8), "save_dirs": [str(tmp_path / f"b{i}") for i in range(3)],
               "request_model_name": "igen", "return_objects": [True, True, True]}
        RM.check_inputdict(req)
        RM.write_framed(b, req)
        b.settimeout(60)
        out = RM.read_framed(b)
        pf = out["pred_frames"]
        assert list(out["save_dirs"]) == req["save_dirs"]
        assert len(pf) == 3                               # Batcher: list of per-branch arrays; no Batcher: ndarray [3, ...]
        for i in range(3):
            fr = np.asarray(pf[i])
            assert fr.shape == (8, 3, 20, 24) and fr.dtype == np.uint8
            for t in range(8):
                assert int(fr[t].min()) == int(fr[t].max()) == _expected(acts[i], t), (i, t)
    finally:
        handler.running = False
        for w in workers:
            try:
                RM.write_pickled_data(w["w_file"], "DONE")
                w["w_file"].close()
            except Exception:
                pass
        for w in workers:
            try:
                assert w["proc"].wait(timeout=30) == 0
            finally:
                if w["proc"].poll() is None:
                    w["proc"].kill()
        a.close(); b.close()
