"""Drop-in WM-server worker for the SVD-dagger ("igen") world model on the sm_100a kernels.

Same process contract as the reference worker FTsvd/eval_inference.py:269-351 behind
downstream/utils/worker_manager.py:
  * spawned by the manager as `<python> wm_worker.py <flags...> <w_fd>` -- the LAST argv is the write-fd;
  * CLI flags of eval_inference.py:273-294 (+ unknown `--k=v` flags the manager forwards are ignored);
  * tasks arrive on stdin as 4-byte big-endian length + pickle of (client_id, task_id, payload);
    results leave on <w_fd> in the same framing (worker_manager.py:186-231, 660-702);
  * payload / result dicts: check_inputdict / check_outputdict (worker_manager.py:106-147),
    process_input_dict / process_output_dict (downstream/api_models/__init__.py:185-237).
Registration: point COMMON_ARGS[PLATFORM]["genex_worker"] (downstream/utils/workers_cfg.py:23-30) at this
script -- see INTEGRATION.md.

Deliberate deviations from the shipped reference, each fixing a latent defect listed in SURVEY.md section 0:
  - world_model_name is defined ("igen", downstream/vlm.py:31) and a name sliced by the manager's Batcher
    (worker_manager.py:457-458: "i", "g", "e", "n", "" ...) is accepted when it is a contiguous slice of it;
  - RGBA observations are sliced to RGB (the reference pipeline fails on 4-channel PIL images);
  - a per-request DATA error (unreadable cond_rgb.png, malformed image / action arrays) is logged to stderr and answered
    with a well-formed reply that carries an explicit `errors` list (one message per branch; zero frames in
    return_objects mode) -- the manager has no error channel (worker_manager.py:369-379) and merges replies key by key
    with list.extend (:476-487), so a per-branch list survives the Batcher.  Everything else -- a wrong
    request_model_name, a failed check_inputdict, a WiwError, any CUDA error (sticky: every later request would
    silently return black frames) -- re-raises and ends the worker loudly, as the reference worker does; so do
    MAX_CONSECUTIVE_DATA_ERRORS data errors in a row.
"""
from __future__ import annotations

import argparse
import json
import os
import pickle
import queue
import sys
import threading
import traceback
from typing import Callable, List, Optional

import numpy as np
import torch

WORLD_MODEL_NAME = "igen"
MAX_CONSECUTIVE_DATA_ERRORS = 3


class RequestDataError(Exception):
    """The request's own data cannot be served (missing / unreadable image file, malformed arrays).  The only error class a
    worker absorbs; configuration, library and device errors propagate."""
IGENEX_ACTION_IDS = {"forward": 1, "turn_left": 2, "turn_right": 3, "stop": 4, "placeholder": 0}


# ----------------------------------------------------------------------------- wire format
def read_frame(fileobj):
    """4-byte big-endian length + pickle (worker_manager.py:218-231)."""
    def exactly(n):
        buf = b""
        while len(buf) < n:
            chunk = fileobj.read(n - len(buf))
            if not chunk:
                raise EOFError(f"Expected {n} bytes, got {len(buf)} before EOF.")
            buf += chunk
        return buf
    n = int.from_bytes(exactly(4), "big")
    if n > 10 ** 14:
        raise ValueError(f"Data length {n} is too large.")
    return pickle.loads(exactly(n))


def write_frame(fileobj, obj) -> None:
    """worker_manager.py:234-239."""
    data = pickle.dumps(obj, protocol=pickle.HIGHEST_PROTOCOL)
    fileobj.write(len(data).to_bytes(4, byteorder="big"))
    fileobj.write(data)
    fileobj.flush()


def check_inputdict(d: dict) -> None:
    """worker_manager.py:106-134 (world_model server type)."""
    assert isinstance(d, dict)
    missing = [k for k in ("b_action", "save_dirs", "request_model_name") if k not in d]
    if missing:
        raise KeyError(f"Missing required keys: {missing}.")
    if "b_image" in d:
        assert isinstance(d["b_image"], np.ndarray) and d["b_image"].dtype == np.uint8
    ba = d["b_action"]
    assert isinstance(ba, list) or (isinstance(ba, np.ndarray) and ba.dtype == np.int64)
    assert isinstance(d["save_dirs"], list) and all(isinstance(s, str) for s in d["save_dirs"])
    if "return_objects" in d:
        assert isinstance(d["return_objects"], list) and all(isinstance(b, bool) for b in d["return_objects"])


def check_outputdict(d: dict) -> None:
    """worker_manager.py:138-147."""
    pf = d.get("pred_frames", None)
    assert pf is None or (isinstance(pf, np.ndarray) and pf.dtype == np.uint8)
    assert "video_tensors" not in d
    assert isinstance(d["save_dirs"], list)


def worker_main(pipe_fd: int, task_fn: Callable[[dict], dict], stdin=None) -> None:
    """Same loop as worker_manager.worker_main (:660-702): a reader thread enqueues frames from stdin, the main
    thread runs task_fn strictly serially and writes (client_id, task_id, result) to pipe_fd; "DONE" stops it."""
    pipe_file = os.fdopen(pipe_fd, "wb", buffering=2048 * 1024)
    src = stdin if stdin is not None else sys.stdin.buffer
    task_q: "queue.Queue" = queue.Queue()
    stop = threading.Event()

    def reader():
        while not stop.is_set():
            try:
                item = read_frame(src)
            except EOFError:
                break
            except Exception as e:          # unrecoverable read error
                task_q.put(("ERROR", -1, {"read_error": str(e)}))
                break
            if isinstance(item, str) and item == "DONE":
                break
            task_q.put(item)
        stop.set()

    threading.Thread(target=reader, daemon=True).start()
    while True:
        try:
            client_id, task_id, payload = task_q.get(timeout=0.1)
        except queue.Empty:
            if stop.is_set():
                break
            continue
        if client_id == "ERROR":
            print(f"[worker] receiver error: {payload}")
            break
        print(f"[worker] Processing task (client={client_id}, task={task_id}).  Remaining queue size: {task_q.qsize()}")
        result = task_fn(payload)
        check_outputdict(result)
        write_frame(pipe_file, (client_id, task_id, result))
    pipe_file.close()


# ----------------------------------------------------------------------------- request helpers
def action_ids_idx_encode(actions: np.ndarray) -> torch.Tensor:
    """utils/svd_utils.py:594-632: A[b,i,j] = act[b,j] for j <= i else 0, column 0 forced to 'stop' (4).  int64 [b,L,L]."""
    actions = np.asarray(actions)
    assert actions.ndim == 2, f"Expected actions to be 2D (batch_size, sequence_length), got {actions.shape}"
    b, L = actions.shape
    a = actions.astype(np.int64).copy()
    a[:, 0] = IGENEX_ACTION_IDS["stop"]
    out = np.zeros((b, L, L), dtype=np.int64)
    for i in range(L):
        out[:, i, : i + 1] = a[:, : i + 1]
    return torch.from_numpy(out)


def name_matches(request_name, deployed: str) -> bool:
    """api_models/__init__.py:189 with the Batcher string-slicing defect tolerated: the manager's
    Batcher slices EVERY request value with v[start:start+bs] (worker_manager.py:457-458), so sub-batch k of an
    "igen" request carries request_model_name == "igen"[k*bs:(k+1)*bs] -- "i", "g", "e", "n", then "" for k >= 4
    (found by running the reference's own manager against this worker, tests/test_reference_manager_interop_cpu.py).
    Accepted: the full name or any contiguous slice of it (including the empty tail slice)."""
    return isinstance(request_name, str) and request_name in deployed


def images_to_tensor(frames_u8: np.ndarray, save_size) -> torch.Tensor:
    """api_models/__init__.py:113-166 (center_crop=False): per frame PIL BICUBIC resize to (w_out, h_out), to_tensor.
    frames_u8 [b, T, 3, H, W] uint8 (what the pipeline's PIL frames hold) -> float32 [b, T, 3, h_out, w_out] in [0,1]."""
    from PIL import Image
    b, T = frames_u8.shape[:2]
    w_out, h_out = save_size
    out = np.empty((b, T, 3, h_out, w_out), dtype=np.float32)
    for i in range(b):
        for t in range(T):
            img = Image.fromarray(np.ascontiguousarray(frames_u8[i, t].transpose(1, 2, 0)))
            img = img.resize((w_out, h_out), Image.BICUBIC)
            out[i, t] = np.asarray(img, dtype=np.float32).transpose(2, 0, 1) / 255.0
    return torch.from_numpy(out)


def save_predict(video: np.ndarray, b_action, save_dirs: List[str]) -> None:
    """downstream/utils/saver.py:221-274: <save_dir>/<idx>.jpg per frame + action_seq.json."""
    from PIL import Image
    for clip, acts, folder in zip(video, b_action, save_dirs):
        os.makedirs(folder, exist_ok=True)
        for i, fr in enumerate(clip):
            Image.fromarray((np.clip(fr, 0, 1) * 255).astype(np.uint8).transpose(1, 2, 0)).save(os.path.join(folder, f"{i}.jpg"))
        with open(os.path.join(folder, "action_seq.json"), "w") as f:
            json.dump(acts.tolist() if hasattr(acts, "tolist") else acts, f, indent=2, ensure_ascii=False)


def save_predict_u8(video_u8: np.ndarray, b_action, save_dirs: List[str]) -> None:
    """save_predict for frames that already went through the reference's resize + double quantisation (uint8 [b,T,3,H,W])."""
    from PIL import Image
    for clip, acts, folder in zip(video_u8, b_action, save_dirs):
        os.makedirs(folder, exist_ok=True)
        for i, fr in enumerate(clip):
            Image.fromarray(np.ascontiguousarray(fr.transpose(1, 2, 0))).save(os.path.join(folder, f"{i}.jpg"))
        with open(os.path.join(folder, "action_seq.json"), "w") as f:
            json.dump(acts.tolist() if hasattr(acts, "tolist") else acts, f, indent=2, ensure_ascii=False)


# ----------------------------------------------------------------------------- the model holder
class Navigator:
    """Counterpart of FTsvd/eval_inference.py:85-266: owns the pipeline and a persistent, seeded generator."""

    def __init__(self, args, rollout, image_encoder, vae_encoder_sd, device):
        self.image_width, self.image_height = args.width, args.height
        self.num_frames = args.num_frames
        self.rollout = rollout
        # CLIP ViT image encoder on the sm_100a kernels (wiw_b200/clip.py); `image_encoder` is either a ready ClipVisionEngine or
        # the transformers module holding the checkpoint (its state dict is re-laid-out once, the module itself is not called)
        from .clip import ClipVisionEngine
        self.image_encoder = image_encoder if isinstance(image_encoder, ClipVisionEngine) else ClipVisionEngine.from_module(image_encoder, device)
        self.vae_encoder_sd = vae_encoder_sd
        self.device = torch.device(device)
        self.generator = torch.Generator(device=self.device).manual_seed(1)        # eval_inference.py:97
        self.num_inference_steps = getattr(args, "num_inference_steps", 30)        # eval_inference.py:255 hard-codes 30
        self.vae_enc = None             # optional engine.VAEEncoderEngine (--own_vae_encoder); default: PyTorch fp32 path like the reference
        self.noise_aug_strength = 0.02

    @torch.no_grad()
    def encode_observation(self, images_u8: np.ndarray, gens=None):
        """Conditioning of one request (pipeline :501-535): uint8 observations [b, 3|4, Hin, Win] of ANY size ->
        (CLIP image embedding [b, 1, D], VAE latents of the noise-augmented frame fp16 [b, 4, H/8, W/8]).
        The reference feeds the observation AS SENT (384x384 / 480x480 from the solvers) to _encode_image, whose antialiased
        resize goes straight to 224x224 (pipeline :192-200), and separately resizes the PIL image to (W, H) with LANCZOS in
        uint8 for the VAE branch (video_processor.preprocess, DP/image_processor.py:501; pipeline :520).
        gens: one torch.Generator per branch (sharded serving) or None for the worker's persistent generator."""
        from . import frontend
        dev = self.device
        H, W = self.image_height, self.image_width
        obs = torch.from_numpy(np.ascontiguousarray(images_u8[:, :3])).to(dev)               # uint8 [b, 3, Hin, Win]
        emb = self.image_encoder(obs)                                                         # [b, 1, D]: resize -> normalise -> ViT, wiw_b200/clip.py
        if tuple(obs.shape[-2:]) != (H, W):
            key = (int(obs.shape[-2]), int(obs.shape[-1]), H, W)
            if getattr(self, "_obs_resizer_key", None) != key:
                from .resize import GpuResizer
                self._obs_resizer, self._obs_resizer_key = GpuResizer(*key, device=dev, requant=False, filt="lanczos"), key
            obs = self._obs_resizer(obs)
        pre = (obs.float() / 255.0) * 2.0 - 1.0                                              # video_processor.preprocess: normalize
        if gens is not None:
            # sharded serving: one generator state per branch, so a branch's noise does not depend on which rank rolls it out
            noise = torch.stack([torch.randn(pre.shape[1:], generator=g, device=dev, dtype=pre.dtype) for g in gens])
        else:
            noise = torch.randn(pre.shape, generator=self.generator, device=dev, dtype=pre.dtype)   # pipeline :521-523
        if self.vae_enc is not None:
            image_latents = self.vae_enc.encode_mode(pre + self.noise_aug_strength * noise)
        else:
            image_latents = frontend.vae_encode_mode(self.vae_encoder_sd, pre + self.noise_aug_strength * noise).to(torch.float16)
        return emb, image_latents

    @torch.no_grad()
    def inference(self, actions: torch.Tensor, images_u8: np.ndarray, out_width=None, out_height=None,
                  latents: Optional[torch.Tensor] = None, on_device: bool = False, seeds: Optional[List[int]] = None,
                  cfg_split_group=None):
        """actions int64/float [b, F, F] (action_ids_idx_encode); images_u8 uint8 [b, 3, H, W].
        Returns uint8 frames [b, F, 3, H, W] (the pipeline's PIL frames).  Branches run one after another
        (request batch 1 per UNet call, as the reference deployment does)."""
        from . import frontend, ops
        from .pipeline import sample_latent_noise_
        from .scheduler import EulerSchedule
        dev = self.device
        b = actions.shape[0]
        F, H, W = self.num_frames, self.image_height, self.image_width
        sched = EulerSchedule(self.num_inference_steps)
        out = torch.empty((b, F, 3, H, W), dtype=torch.uint8, device=dev)
        gens = [torch.Generator(device=dev).manual_seed(int(sd)) for sd in seeds] if seeds is not None else None
        emb, image_latents = self.encode_observation(images_u8, gens)
        if latents is None and seeds is not None:
            raw = torch.stack([torch.randn((F, 4, H // 8, W // 8), generator=g, device=dev, dtype=torch.float16) for g in gens])
            diag = torch.diagonal(actions, dim1=-2, dim2=-1)
            for i in range(b):
                sample_latent_noise_(raw[i], diag[i])
        elif latents is None:                                                                # prepare_latents :325-362
            raw = torch.randn((b, F, 4, H // 8, W // 8), generator=self.generator, device=dev, dtype=torch.float16)
            diag = torch.diagonal(actions, dim1=-2, dim2=-1)                                 # decode_action_seq_frames :826-844
            for i in range(b):
                sample_latent_noise_(raw[i], diag[i])
        else:
            raw = latents.to(dev, torch.float16)
        raw = raw * sched.init_noise_sigma                                                   # :361
        for i in range(b):
            lat = self.rollout.denoise(raw[i], image_latents[i], emb[i], actions[i].float(), num_steps=self.num_inference_steps,
                                       noise_aug_strength=self.noise_aug_strength, cfg_split_group=cfg_split_group)
            out[i] = self.rollout.decode(lat)
        return out if on_device else out.cpu().numpy()


class WorldModelWorker:
    """task_fn of the worker: request dict -> response dict (eval_inference.py:313-349)."""

    def __init__(self, args, navigator: Navigator):
        self.args, self.nav = args, navigator

    def resized_frames(self, frames: torch.Tensor) -> torch.Tensor:
        """uint8 [b, F, 3, H, W] on the device -> uint8 [b, F, 3, H_out, W_out]: PIL BICUBIC resize + the reference's double
        quantisation, byte-exact, on the GPU (wiw_resize_bicubic_u8)."""
        key = (frames.shape[-2], frames.shape[-1], self.args.out_height, self.args.out_width)
        if getattr(self, "_resizer_key", None) != key:
            from .resize import GpuResizer
            self._resizer, self._resizer_key = GpuResizer(*key, device=frames.device), key
        return self._resizer(frames)

    def _validate(self, input_dict: dict) -> None:
        """Configuration-level checks: a failure here is not a data error and ends the worker (the reference asserts too,
        api_models/__init__.py:189, worker_manager.py:106-134)."""
        check_inputdict(input_dict)
        if not name_matches(input_dict["request_model_name"], WORLD_MODEL_NAME):
            raise ValueError(f"request_model_name: {input_dict['request_model_name']} does not match deployed world_model_name: "
                             f"{WORLD_MODEL_NAME}")

    @staticmethod
    def _request_arrays(input_dict: dict, save_dirs: List[str]):
        """b_action / b_image of a validated request; anything wrong with the DATA raises RequestDataError."""
        try:
            b_action = np.asarray(input_dict["b_action"])
            b_image = input_dict.get("b_image")
            if b_image is None:
                from PIL import Image
                b_image = np.stack([np.asarray(Image.open(os.path.join(d, "cond_rgb.png")).convert("RGB")).transpose(2, 0, 1)
                                    for d in save_dirs])
        except (OSError, ValueError) as e:
            raise RequestDataError(f"cannot read the request's observation / actions: {type(e).__name__}: {e}") from e
        if b_image.ndim != 4 or b_image.dtype != np.uint8 or b_image.shape[1] not in (3, 4):
            raise RequestDataError(f"b_image should be uint8 B C H W with 3 or 4 channels, got {b_image.dtype} {b_image.shape}")
        if b_action.ndim != 2 or len(b_action) != len(save_dirs) or len(b_image) != len(save_dirs):
            raise RequestDataError(f"b_action {b_action.shape} / b_image {b_image.shape} do not match {len(save_dirs)} save_dirs")
        return b_action, b_image

    def _data_error_reply(self, err: Exception, save_dirs: List[str], return_objects: bool) -> dict:
        """Well-formed reply for an absorbed data error, with an explicit marker: `errors` (one message per branch)."""
        self._consecutive_data_errors = getattr(self, "_consecutive_data_errors", 0) + 1
        print(f"[igenex worker] REQUEST FAILED ({self._consecutive_data_errors} in a row): {err}", file=sys.stderr, flush=True)
        if self._consecutive_data_errors >= MAX_CONSECUTIVE_DATA_ERRORS:
            raise RuntimeError(f"{self._consecutive_data_errors} consecutive failed requests; last: {err}") from err
        n = len(save_dirs)
        out = {"save_dirs": save_dirs, "errors": [str(err)] * n}
        if return_objects:
            out["pred_frames"] = np.zeros((n, self.args.num_frames, 3, self.args.out_height, self.args.out_width), dtype=np.uint8)
        return out

    def do_some_tasks(self, input_dict: dict) -> dict:
        self._validate(input_dict)                                  # raises: not a per-request data problem
        save_dirs = list(input_dict["save_dirs"])
        return_objects = bool(input_dict.get("return_objects"))
        try:
            b_action, b_image = self._request_arrays(input_dict, save_dirs)
        except RequestDataError as e:
            return self._data_error_reply(e, save_dirs, return_objects)
        # from here on every failure is the library's or the device's (WiwError, CUDA): it propagates and ends the worker
        action_ids = action_ids_idx_encode(b_action)
        frames = self.nav.inference(action_ids, b_image, self.args.out_width, self.args.out_height, on_device=True)
        small = self.resized_frames(frames).cpu().numpy()           # only the 480x480 uint8 result crosses PCIe
        self._consecutive_data_errors = 0
        if return_objects:
            return {"pred_frames": small, "save_dirs": save_dirs}
        # files mode (saver.py:221-274): same device-side resize + requantisation, the JPEGs hold exactly the bytes the reference's
        # images_to_tensor -> save_predict chain would write
        save_predict_u8(small, b_action, save_dirs)
        return {"save_dirs": save_dirs}


class ShardedWorldModelWorker(WorldModelWorker):
    """One worker that owns R GPUs of a node (BASELINE configs[2]): rank 0 is the process the manager
    spawned (stdin / write-fd); ranks 1..R-1 are helper processes it started.  Per request rank 0 broadcasts the payload
    (control plane, gloo), rank r rolls out branches r, r+R, ... with no per-step communication, and ONE all-gather of the
    resized uint8 frames (NCCL over NVLink on the GPU box) returns the result to rank 0, which answers the manager.
    Launch with `--gpus R` and run the manager with a batch size >= the request size so it does not pre-split the batch."""

    def __init__(self, args, navigator, rank: int, world: int):
        super().__init__(args, navigator)
        self.rank, self.world, self.serial = rank, world, 0
        # --cfg_split: a request with at most R/2 branches (solver_AR sends 2) runs every branch on a PAIR of GPUs, one row of the
        # classifier-free-guidance pair each (Rollout.denoise cfg_split_group): ~1.85x lower latency instead of idle GPUs
        self.pair_group = None
        if getattr(args, "cfg_split", False) and world >= 2 and world % 2 == 0:
            import torch.distributed as dist
            for k in range(world // 2):                   # collective: every rank creates every pair group
                grp = dist.new_group([2 * k, 2 * k + 1])
                if rank // 2 == k:
                    self.pair_group = grp

    def _use_pairs(self, b: int) -> bool:
        return self.pair_group is not None and 2 * b <= self.world

    def _local(self, input_dict: dict):
        from .sharding import branch_indices, pair_branch_indices
        b = len(input_dict["save_dirs"])
        idx = pair_branch_indices(b, self.rank, self.world) if self._use_pairs(b) else branch_indices(b, self.rank, self.world)
        local = {k: ([v[i] for i in idx] if isinstance(v, list) else (v[idx] if isinstance(v, np.ndarray) else v))
                 for k, v in input_dict.items()}
        return b, idx, local

    def serve_request(self, input_dict: Optional[dict]) -> Optional[dict]:
        """Collective: every rank calls it once per request (rank 0 with the payload, the helpers with None)."""
        import torch.distributed as dist
        from .sharding import gather_frames, gather_frames_pairs
        box = [input_dict]
        dist.broadcast_object_list(box, src=0)
        req = box[0]
        if req is None:                                   # shutdown
            return None
        self.serial += 1
        F, Ho, Wo = self.args.num_frames, self.args.out_height, self.args.out_width
        dev = self.nav.device
        # outcome per rank: 0 ok, 1 data error (absorbed: explicit `errors` reply), 2 anything else (fatal on EVERY rank)
        status, err, b, idx, loc, acts, small = 0, None, 0, [], {}, None, None
        return_objects = bool(req.get("return_objects")) if isinstance(req, dict) else False
        try:
            self._validate(req)
            b, idx, loc = self._local(req)
            if idx:
                acts, b_image = self._request_arrays(loc, loc["save_dirs"])
                seeds = [1 + 1000003 * self.serial + i for i in idx]          # per (request, branch): independent of the rank count
                frames = self.nav.inference(action_ids_idx_encode(acts), b_image, Wo, Ho, on_device=True, seeds=seeds,
                                            cfg_split_group=self.pair_group if self._use_pairs(b) else None)
                small = self.resized_frames(frames)
            else:
                acts = np.zeros((0, F), np.int64)
                small = torch.zeros((0, F, 3, Ho, Wo), dtype=torch.uint8, device=dev)
        except RequestDataError as e:
            status, err = 1, e
        except Exception as e:
            traceback.print_exc()
            status, err = 2, e
        # one rank failing must not leave the others waiting in the all-gather: agree on the outcome first (control plane, gloo)
        flag = torch.tensor([status], dtype=torch.int32)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        worst = int(flag.item())
        if worst == 2:                                    # configuration / library / device error somewhere: every rank stops
            raise RuntimeError(f"rank {self.rank}: request {self.serial} failed fatally on some rank") from err
        if worst == 1:
            msg = err if err is not None else RequestDataError("the request's data could not be served on another rank")
            return self._data_error_reply(msg, list(req.get("save_dirs", [])), return_objects)
        self._consecutive_data_errors = 0
        if return_objects:
            out = (gather_frames_pairs if self._use_pairs(b) else gather_frames)(small, b)      # the one data-path collective
            return {"pred_frames": out.cpu().numpy(), "save_dirs": list(req["save_dirs"])} if self.rank == 0 else {}
        if idx and not (self._use_pairs(b) and self.rank % 2):             # files mode: every rank (pair leader) writes its own branches
            save_predict_u8(small.cpu().numpy(), acts, loc["save_dirs"])
        dist.barrier()
        return {"save_dirs": list(req["save_dirs"])} if self.rank == 0 else {}

    def do_some_tasks(self, input_dict: dict) -> dict:            # rank 0's task_fn
        return self.serve_request(input_dict)

    def helper_loop(self) -> None:                                # ranks 1..R-1
        while self.serve_request(None) is not None:
            pass


# ----------------------------------------------------------------------------- construction
def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(description="SVD-dagger world-model worker (B200 kernels)")
    p.add_argument("--width", type=int, default=1024)
    p.add_argument("--height", type=int, default=576)
    p.add_argument("--out_width", type=int, default=480)
    p.add_argument("--out_height", type=int, default=480)
    p.add_argument("--log_dir", type=str, default="downstream/logs")
    p.add_argument("--exp_id", type=str, default="03.15_usegenex")
    p.add_argument("--num_frames", type=int, default=14)
    p.add_argument("--num_past_obs", type=int, default=1)
    p.add_argument("--task_type", type=str, default="navigation", choices=["navigation", "manipulation", "freetext"])
    p.add_argument("--action_strategy", type=str, default="micro_cond")
    p.add_argument("--action_input_channel", type=int, default=14)
    p.add_argument("--device", type=str, default="cuda:0")
    p.add_argument("--unet_path", type=str, default="")
    p.add_argument("--svd_path", type=str, default="")
    p.add_argument("--weight_dtype", type=str, default="float16")
    # additions (not in the reference): synthetic weights when no checkpoint is reachable, step count, tiny geometry
    p.add_argument("--random_init", action="store_true")
    p.add_argument("--num_inference_steps", type=int, default=30)
    p.add_argument("--tiny", action="store_true")
    # sharded serving (BASELINE configs[2]): this worker owns `--gpus` GPUs of the node; --rank / --master_port are set by
    # rank 0 when it starts its helper processes
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--rank", type=int, default=0)
    p.add_argument("--master_port", type=int, default=0)
    p.add_argument("--own_vae_encoder", action="store_true",
                   help="run the conditioning frame's VAE encode on the sm_100a kernels in fp16 (opt-in; the reference and the "
                        "default path run it in fp32 on PyTorch library kernels)")
    p.add_argument("--cfg_split", action="store_true",
                   help="with --gpus R: requests of at most R/2 branches run each branch on a pair of GPUs (one CFG row each)")
    return p


def build_worker(args) -> WorldModelWorker:
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from . import weights as W
    from .config import TINY_UNET, TINY_VAE, UNetConfig, VAEConfig
    from .engine import UNetEngine, VAEDecoderEngine
    from .frontend import vae_encoder_param_spec
    from .lib import require_device
    from .pipeline import Rollout
    require_device()
    if args.num_past_obs != 1 or args.action_strategy != "micro_cond" or args.task_type != "navigation":
        raise ValueError("this worker serves the deployed setting: micro_cond, navigation, num_past_obs == 1")
    dev = torch.device(args.device)
    torch.cuda.set_device(dev)
    if args.tiny:
        ucfg, vcfg = TINY_UNET, TINY_VAE
        ucfg = UNetConfig(**{**ucfg.__dict__, "num_frames": args.num_frames, "action_input_channel": args.action_input_channel})
    else:
        ucfg = UNetConfig(num_frames=args.num_frames, action_input_channel=args.action_input_channel)
        vcfg = VAEConfig()
    if args.random_init or not args.unet_path:
        usd = W.random_init_on_device(W.unet_param_spec(ucfg), 0, dev)
        vsd = W.random_init_on_device(W.vae_decoder_param_spec(vcfg), 1, dev)
        esd = {k: v.float() for k, v in W.random_init_on_device(
            vae_encoder_param_spec(vcfg.block_out_channels, vcfg.layers_per_block, vcfg.latent_channels), 2, dev).items()}
        torch.manual_seed(3)
        clip_cfg = (CLIPVisionConfig(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                                     patch_size=32, projection_dim=ucfg.cross_attention_dim) if args.tiny else
                    CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                                     patch_size=14, projection_dim=1024))
        clip = CLIPVisionModelWithProjection(clip_cfg).to(dev, torch.float16).eval()
    else:
        usd = W.load_safetensors_dir(args.unet_path if os.path.basename(args.unet_path.rstrip("/")) == "unet"
                                     else os.path.join(args.unet_path, "unet"))
        W.check_state_dict(usd, W.unet_param_spec(ucfg))
        full_vae = W.load_safetensors_dir(os.path.join(args.svd_path, "vae"))
        vsd = {k[len("decoder."):]: v for k, v in full_vae.items() if k.startswith("decoder.")}
        esd = {k: v.float().to(dev) for k, v in full_vae.items() if not k.startswith("decoder.")}
        clip = CLIPVisionModelWithProjection.from_pretrained(os.path.join(args.svd_path, "image_encoder"),
                                                             torch_dtype=torch.float16).to(dev).eval()
    rollout = Rollout(UNetEngine(ucfg, usd, dev), VAEDecoderEngine(vcfg, vsd, dev), dev)
    del usd, vsd
    nav = Navigator(args, rollout, clip, esd, dev)
    if getattr(args, "own_vae_encoder", False):
        from .engine import VAEEncoderEngine
        nav.vae_enc = VAEEncoderEngine(vcfg, esd, dev)
    if getattr(args, "gpus", 1) > 1:
        return ShardedWorldModelWorker(args, nav, args.rank, args.gpus)
    return WorldModelWorker(args, nav)


def _free_port() -> int:
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def main(argv: List[str]) -> None:
    pipe_fd = int(argv[-1])                                    # eval_inference.py:271
    args, _unknown = build_parser().parse_known_args(argv[:-1])
    print(f"[igenex worker] All Args:\n {args}")
    helpers = []
    if args.gpus > 1:
        import subprocess
        import torch.distributed as dist
        if args.rank == 0:                                     # the process the manager spawned: start ranks 1..R-1
            args.master_port = args.master_port or _free_port()
            for r in range(1, args.gpus):                     # same flags; argparse keeps the LAST occurrence of the overrides
                helpers.append(subprocess.Popen([sys.executable, os.path.abspath(__file__)] + list(argv[:-1]) +
                                                [f"--rank={r}", f"--master_port={args.master_port}", f"--device=cuda:{r}", "-1"],
                                                stdin=subprocess.DEVNULL))
        args.device = f"cuda:{args.rank}"
        torch.cuda.set_device(args.rank)
        dist.init_process_group("cpu:gloo,cuda:nccl", init_method=f"tcp://127.0.0.1:{args.master_port}", rank=args.rank,
                                world_size=args.gpus, device_id=torch.device(args.device))
    worker = build_worker(args)
    print(f"[igenex worker] Initialized navigator with device {args.device}.")
    if args.gpus > 1 and args.rank > 0:
        worker.helper_loop()
        return
    worker_main(pipe_fd, worker.do_some_tasks)
    if args.gpus > 1:
        worker.serve_request(None)                             # tell the helpers to exit
        for h in helpers:
            h.wait(timeout=120)


if __name__ == "__main__":
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if here not in sys.path:
        sys.path.insert(0, here)
    import wiw_b200  # noqa: F401  (registers the package so the relative imports above resolve)
    from wiw_b200.wm_worker import main as _main
    _main(sys.argv[1:])
