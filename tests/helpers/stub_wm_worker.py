"""Test helper (CPU only): the REAL wiw_b200 worker process loop and request handling -- wm_worker.worker_main +
WorldModelWorker.do_some_tasks -- around a fake Navigator that needs no GPU.  Spawned by
tests/test_reference_manager_interop_cpu.py through the reference's own manager code, exactly as a genex worker
would be (`<python> <script> <flags...> <w_fd>`).  Frame content encodes the request so the test can check routing:
pixel value of frame t of a branch = (sum(actions) * 7 + t) % 256 (+100 when the branch ran as a CFG pair)."""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, REPO)
import wiw_b200  # noqa: F401,E402
from wiw_b200 import wm_worker as W  # noqa: E402


class FakeNavigator:
    def __init__(self, args):
        self.args = args
        self.device = torch.device("cpu")

    def inference(self, actions, images_u8, out_width=None, out_height=None, latents=None, on_device=False, seeds=None,
                  cfg_split_group=None):
        b, F = actions.shape[0], self.args.num_frames
        assert images_u8.shape[1] in (3, 4) and images_u8.dtype == np.uint8
        acts = torch.diagonal(actions, dim1=-2, dim2=-1)
        out = torch.empty((b, F, 3, self.args.height, self.args.width), dtype=torch.uint8)
        for i in range(b):
            if int(acts[i].sum()) == 77:
                raise RuntimeError("stub: injected failure for this branch")
            for t in range(F):
                out[i, t] = int((int(acts[i].sum()) * 7 + t + (100 if cfg_split_group is not None else 0)) % 256)
        return out


def cpu_resizer(args):
    """The reference's own post-processing (api_models/__init__.py:113-166,229) on the CPU: PIL bicubic + double quantisation."""
    def run(frames):
        v = W.images_to_tensor(frames.numpy(), save_size=(args.out_width, args.out_height)).numpy()
        return torch.from_numpy((np.clip(v, 0, 1) * 255).astype(np.uint8))
    return run


if __name__ == "__main__":
    pipe_fd = int(sys.argv[-1])
    args, _ = W.build_parser().parse_known_args(sys.argv[1:-1])
    worker = W.WorldModelWorker(args, FakeNavigator(args))
    worker._resizer_key = (args.height, args.width, args.out_height, args.out_width)
    worker._resizer = cpu_resizer(args)
    W.worker_main(pipe_fd, worker.do_some_tasks)
