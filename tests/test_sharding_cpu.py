"""World-size-2 gloo test of the branch sharding + single all-gather host logic (the N>1 path of bench.py / the worker)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import wiw_b200  # noqa: F401
from wiw_b200 import sharding


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(rank, world, port, b, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    idx = sharding.branch_indices(b, rank, world)
    local = torch.stack([torch.full((2, 3, 4, 5), i, dtype=torch.uint8) for i in idx]) if idx else torch.zeros(0, 2, 3, 4, 5, dtype=torch.uint8)
    out = sharding.gather_frames(local, b)
    ok = out.shape == (b, 2, 3, 4, 5) and all(bool((out[i] == i).all()) for i in range(b))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_branch_sharding_and_gather_world2():
    assert sharding.branch_indices(5, 1, 2) == [1, 3] and sharding.max_local(5, 2) == 3
    ctx = mp.get_context("spawn")
    for b in (5, 2, 1):
        q, port = ctx.Queue(), _free_port()
        ps = [ctx.Process(target=_run, args=(r, 2, port, b, q)) for r in range(2)]
        [p.start() for p in ps]
        res = [q.get(timeout=120) for _ in ps]
        [p.join(60) for p in ps]
        assert all(ok for _, ok in res), res
