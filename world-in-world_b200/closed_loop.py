"""Closed-loop serving harness for BASELINE configs[3]: a heuristic solver that plans with the world model over a short horizon,
talking to the WM worker through a manager exactly as the reference's solvers do.

Two small pieces, both mirrors of the reference's CALLER side of the hot path so that the worker can be
driven end to end without the reference tree (which does not exist on the GPU box):

  MiniManager      the slice of downstream/utils/worker_manager.py a single igenex worker needs: spawn the worker as
                   `<python> <script> <flags...> <w_fd>` with stdin = PIPE (create_workers :303-346), accept solver connections on a
                   TCP socket, split every request into sub-batches of `batch_size` by slicing EVERY value with v[k:k+bs] -- the
                   request_model_name string included (Batcher.split_batch :448-469) --, frame them to the worker's stdin as
                   (client_id, task_id, payload) (:186-239), read the framed results from the worker's pipe and recompose them key
                   by key with list.extend in task order (_recompose_batch :476-487).  The reference's fixed sleeps (0.06 s per
                   dispatched sub-batch :570, 0.05 s client poll :548, 2 s after each spawn :335) are NOT reproduced: they are
                   overheads of that manager, not part of the contract.
  HeuristicSolver  the request side of downstream/solver_base.py: `send_batch_to_server` (:645-673) builds
                   {b_action int64[b,14], b_image uint8[b,C,H,W], save_dirs (absolute), request_model_name, return_objects} and
                   reads back {pred_frames uint8[b,14,3,H,W], save_dirs}; action rows are `[stop] + prior + [forward] * rest`
                   (_prepare_common_data :474-500).  The policy is a deterministic heuristic (no VLM, no simulator): every step it
                   proposes `branches` candidate action prefixes, scores the LAST predicted frame of each branch and commits to the
                   best one; with no simulator in the loop the committed branch's last predicted frame becomes the next
                   observation, so every request depends on the previous answer (a genuinely closed loop).

Nothing here touches the GPU: it is host-side plumbing around the worker process.
"""
from __future__ import annotations

import os
import pickle
import socket
import subprocess
import sys
import threading
import time
from typing import Dict, List, Optional, Sequence

import numpy as np

IGENEX_ACTION_IDS = {"forward": 1, "turn_left": 2, "turn_right": 3, "stop": 4, "placeholder": 0}


# ----------------------------------------------------------------------------- framing (worker_manager.py:150-239)
def _recv_exactly(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise EOFError(f"Expected {n} bytes, got {len(buf)} before EOF.")
        buf += chunk
    return bytes(buf)


def sock_write_framed(sock: socket.socket, obj) -> None:
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    sock.sendall(len(data).to_bytes(4, "big") + data)


def sock_read_framed(sock: socket.socket):
    n = int.from_bytes(_recv_exactly(sock, 4), "big")
    if n > 10 ** 14:
        raise ValueError(f"Data length {n} is too large.")
    return pickle.loads(_recv_exactly(sock, n))


def split_batch(tasks: dict, bs: int) -> List[dict]:
    """Batcher.split_batch (worker_manager.py:448-469): slice every value, strings included."""
    n = len(next(iter(tasks.values())))
    assert n > 0, f"Data length must be > 0, but got {n}"
    return [{k: v[s:s + bs] for k, v in tasks.items()} for s in range(0, n, bs)]


def recompose(results: Sequence[dict]) -> dict:
    """Batcher._recompose_batch (worker_manager.py:476-487): per key, extend a list in task order."""
    out: Dict[str, list] = {}
    for item in results:
        for k, v in item.items():
            out.setdefault(k, []).extend(v)
    return out


# ----------------------------------------------------------------------------- manager
class MiniManager:
    def __init__(self, worker_cmd: List[str], batch_size: int = 1, host: str = "127.0.0.1", port: int = 0, cwd: Optional[str] = None):
        self.batch_size = batch_size
        self.r_fd, w_fd = os.pipe()
        self.proc = subprocess.Popen(list(worker_cmd) + [str(w_fd)], stdin=subprocess.PIPE, pass_fds=[w_fd], cwd=cwd)
        os.close(w_fd)
        self.w_file = self.proc.stdin
        self.r_file = os.fdopen(self.r_fd, "rb")
        self.srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind((host, port))
        self.srv.listen(4)
        self.host, self.port = self.srv.getsockname()
        self.lock = threading.Lock()                      # one worker, strictly serial (worker_main :660-702)
        self.next_client = 0
        self.sub_batches = 0
        self.running = True
        self.thread = threading.Thread(target=self._accept_loop, daemon=True)
        self.thread.start()

    def _worker_roundtrip(self, client_id: int, task_id: int, payload: dict) -> dict:
        data = pickle.dumps((client_id, task_id, payload), protocol=pickle.HIGHEST_PROTOCOL)
        self.w_file.write(len(data).to_bytes(4, "big") + data)
        self.w_file.flush()
        hdr = self.r_file.read(4)
        if len(hdr) < 4:
            raise EOFError(f"worker closed its pipe (exit code {self.proc.poll()})")
        cid, tid, result = pickle.loads(self.r_file.read(int.from_bytes(hdr, "big")))
        assert (cid, tid) == (client_id, task_id), f"worker answered ({cid}, {tid}) to ({client_id}, {task_id})"
        return result

    def _serve_client(self, conn: socket.socket, client_id: int) -> None:
        task_id = 0
        try:
            while True:
                try:
                    msg = sock_read_framed(conn)
                except EOFError:
                    break
                if isinstance(msg, str) and msg == "DONE":
                    break
                parts = split_batch(msg, self.batch_size) if self.batch_size > 0 else [msg]
                results = []
                with self.lock:
                    for part in parts:
                        results.append(self._worker_roundtrip(client_id, task_id, part))
                        task_id += 1
                        self.sub_batches += 1
                sock_write_framed(conn, recompose(results) if self.batch_size > 0 else results[0])
        except Exception as e:                              # surface worker death to the client instead of hanging it
            print(f"[mini manager] client {client_id}: {type(e).__name__}: {e}", file=sys.stderr, flush=True)
        finally:
            conn.close()

    def _accept_loop(self) -> None:
        while self.running:
            try:
                conn, _ = self.srv.accept()
            except OSError:
                break
            cid, self.next_client = self.next_client, self.next_client + 1
            threading.Thread(target=self._serve_client, args=(conn, cid), daemon=True).start()

    def stop(self, timeout: float = 120.0) -> int:
        self.running = False
        try:
            self.srv.close()
        except OSError:
            pass
        try:
            data = pickle.dumps("DONE", protocol=pickle.HIGHEST_PROTOCOL)
            self.w_file.write(len(data).to_bytes(4, "big") + data)
            self.w_file.flush()
            self.w_file.close()
        except (OSError, ValueError):
            pass
        try:
            return self.proc.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            return -9


# ----------------------------------------------------------------------------- solver
def frame_score(frame_u8: np.ndarray) -> float:
    """Deterministic stand-in for the VLM's judgement of a predicted view: mean absolute horizontal + vertical gradient of the
    luma (a 'how much structure is in sight' score).  Integer arithmetic on uint8 frames, so it is reproducible bit for bit."""
    y = frame_u8.astype(np.int32).sum(axis=0)                                        # [H, W], 3 x luma
    return float(np.abs(np.diff(y, axis=0)).mean() + np.abs(np.diff(y, axis=1)).mean())


class HeuristicSolver:
    def __init__(self, host: str, port: int, save_root: str, num_frames: int = 14, world_model_name: str = "igen"):
        self.sock = socket.create_connection((host, port))
        self.save_root, self.F, self.name = save_root, num_frames, world_model_name
        self.requests = 0

    def action_rows(self, prefixes: Sequence[Sequence[str]]) -> np.ndarray:
        """_prepare_common_data (solver_base.py:488-495): [stop] + prior action ids + [forward] * remaining."""
        rows = []
        for p in prefixes:
            ids = [IGENEX_ACTION_IDS[a] for a in p]
            assert len(ids) <= self.F - 1
            rows.append([IGENEX_ACTION_IDS["stop"]] + ids + [IGENEX_ACTION_IDS["forward"]] * (self.F - 1 - len(ids)))
        return np.asarray(rows, dtype=np.int64)

    def look_ahead(self, obs_u8: np.ndarray, prefixes: Sequence[Sequence[str]], step: int) -> dict:
        """send_batch_to_server (solver_base.py:645-673) with the same observation for every branch."""
        b = len(prefixes)
        save_dirs = [os.path.abspath(os.path.join(self.save_root, f"step{step}", "igenex", f"action_{i}")) for i in range(b)]
        req = {"b_action": self.action_rows(prefixes), "save_dirs": save_dirs, "request_model_name": self.name,
               "b_image": np.ascontiguousarray(np.broadcast_to(obs_u8, (b,) + obs_u8.shape)), "return_objects": [True] * b}
        sock_write_framed(self.sock, req)
        self.requests += 1
        out = sock_read_framed(self.sock)
        out["pred_frames"] = np.asarray(out["pred_frames"], dtype=np.uint8)         # the Batcher hands back a list of per-branch arrays
        return out

    def run(self, obs_u8: np.ndarray, horizon: int = 4, branches: int = 2) -> dict:
        candidates = [["forward"], ["turn_left", "turn_left"], ["turn_right", "turn_right"], ["turn_left", "forward", "turn_right"]][:branches]
        chosen, scores_log, t_req = [], [], []
        frames = 0
        t0 = time.perf_counter()
        for step in range(horizon):
            t1 = time.perf_counter()
            out = self.look_ahead(obs_u8, candidates, step)
            t_req.append(time.perf_counter() - t1)
            pf = out["pred_frames"]
            if "errors" in out:
                raise RuntimeError(f"world model reported: {out['errors']}")
            assert pf.shape[:3] == (len(candidates), self.F, 3), pf.shape
            scores = [frame_score(pf[i, -1]) for i in range(len(candidates))]
            best = int(np.argmax(scores))
            chosen.append(best)
            scores_log.append(scores)
            frames += pf.shape[0] * pf.shape[1]
            obs_u8 = np.ascontiguousarray(pf[best, -1])                              # closed loop: the prediction is the next observation
        dt = time.perf_counter() - t0
        return {"horizon": horizon, "branches": len(candidates), "requests": self.requests, "frames": frames, "seconds": dt,
                "requests_per_s": horizon / dt, "frames_per_s": frames / dt, "seconds_per_request": t_req, "chosen": chosen,
                "scores": scores_log, "final_obs_shape": list(obs_u8.shape)}

    def close(self) -> None:
        try:
            sock_write_framed(self.sock, "DONE")
        finally:
            self.sock.close()
