"""World-size-2 gloo test of the sharded WM worker (BASELINE configs[2] host logic): rank 0 broadcasts the request, each rank
serves its branches (CPU fake Navigator), one all-gather returns the frames in request order; files mode; shutdown."""
import os
import socket
import sys

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

import wiw_b200  # noqa: F401

HELPERS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "helpers")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _expected(row, t):
    a = np.array(row).copy()
    a[0] = 4
    return (int(a.sum()) * 7 + t) % 256


def _run(rank, world, port, tmp, q):
    sys.path.insert(0, HELPERS)
    import stub_wm_worker as S
    from wiw_b200 import wm_worker as W
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    args = W.build_parser().parse_args(["--width", "64", "--height", "32", "--out_width", "24", "--out_height", "20", "--num_frames", "8",
                                        "--action_input_channel", "8", "--gpus", str(world), "--rank", str(rank)])
    worker = W.ShardedWorldModelWorker(args, S.FakeNavigator(args), rank, world)
    worker._resizer_key = (args.height, args.width, args.out_height, args.out_width)
    worker._resizer = S.cpu_resizer(args)
    if rank > 0:
        try:
            worker.helper_loop()
            q.put((rank, False))                                 # the last request of rank 0's script must end every rank
        except RuntimeError:
            q.put((rank, True))
        dist.destroy_process_group()
        return
    ok = True
    acts = np.array([[4, 1, 2, 2, 1, 3, 1, 1], [1, 3, 3, 1, 1, 1, 2, 0], [2, 2, 2, 2, 1, 1, 1, 1]], dtype=np.int64)
    for b in (3, 1):                                             # b < world: one rank idles but still joins the collectives
        req = {"b_action": acts[:b], "b_image": np.zeros((b, 4, 32, 64), np.uint8), "save_dirs": [os.path.join(tmp, f"o{b}_{i}") for i in range(b)],
               "request_model_name": "igen", "return_objects": [True] * b}
        out = worker.do_some_tasks(req)
        W.check_outputdict(out)
        pf = out["pred_frames"]
        ok &= pf.shape == (b, 8, 3, 20, 24) and out["save_dirs"] == req["save_dirs"]
        ok &= all(int(pf[i, t].min()) == int(pf[i, t].max()) == _expected(acts[i], t) for i in range(b) for t in range(8))
    req = {"b_action": acts, "b_image": np.zeros((3, 3, 32, 64), np.uint8), "save_dirs": [os.path.join(tmp, f"f{i}") for i in range(3)],
           "request_model_name": "igen"}
    out = worker.do_some_tasks(req)                               # files mode: every rank writes its own branches
    ok &= "pred_frames" not in out and all(sorted(os.listdir(d)) == sorted([f"{t}.jpg" for t in range(8)] + ["action_seq.json"])
                                             for d in req["save_dirs"])
    # data error (no b_image, no cond_rgb.png): every rank absorbs it alike -- zero frames + an explicit `errors` list, no hang
    nofile = {k: v for k, v in req.items() if k != "b_image"}
    bad = worker.do_some_tasks({**nofile, "return_objects": [True] * 3})
    ok &= bad["pred_frames"].shape == (3, 8, 3, 20, 24) and int(bad["pred_frames"].max()) == 0 and len(bad["errors"]) == 3
    out = worker.do_some_tasks({**req, "return_objects": [True] * 3})                                   # and the worker keeps serving
    ok &= int(out["pred_frames"][2].max()) == _expected(acts[2], 7) and "errors" not in out
    # a library-side failure on ONE rank (branch 1 lives on rank 1 and raises there only) is fatal on EVERY rank: agreed before
    # the all-gather, so nobody hangs and nobody returns black frames
    acts_bad = acts.copy()
    acts_bad[1] = [4, 73, 0, 0, 0, 0, 0, 0]
    try:
        worker.do_some_tasks({**req, "b_action": acts_bad, "return_objects": [True] * 3})
        ok = False
    except RuntimeError:
        pass
    q.put((rank, bool(ok)))
    dist.destroy_process_group()
    return
    worker.serve_request(None)                                    # shutdown broadcast
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_worker_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_run, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=180) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res), res


def _run_pairs(rank, world, port, tmp, q):
    sys.path.insert(0, HELPERS)
    import stub_wm_worker as S
    from wiw_b200 import wm_worker as W
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    args = W.build_parser().parse_args(["--width", "64", "--height", "32", "--out_width", "24", "--out_height", "20", "--num_frames", "8",
                                        "--action_input_channel", "8", "--gpus", str(world), "--rank", str(rank), "--cfg_split"])
    worker = W.ShardedWorldModelWorker(args, S.FakeNavigator(args), rank, world)
    worker._resizer_key = (args.height, args.width, args.out_height, args.out_width)
    worker._resizer = S.cpu_resizer(args)
    if rank > 0:
        worker.helper_loop()
        q.put((rank, True))
        dist.destroy_process_group()
        return
    ok = worker.pair_group is not None
    acts = np.array([[4, 1, 2, 2, 1, 3, 1, 1], [1, 3, 3, 1, 1, 1, 2, 0], [2, 2, 2, 2, 1, 1, 1, 1]], dtype=np.int64)
    for b, paired in ((2, True), (1, True), (3, False)):            # 2*b <= 4 ranks -> CFG pairs; 3 branches -> one branch per rank
        req = {"b_action": acts[:b], "b_image": np.zeros((b, 3, 32, 64), np.uint8), "save_dirs": [os.path.join(tmp, f"p{b}_{i}") for i in range(b)],
               "request_model_name": "igen", "return_objects": [True] * b}
        pf = worker.do_some_tasks(req)["pred_frames"]
        ok &= pf.shape == (b, 8, 3, 20, 24)
        bump = 100 if paired else 0
        ok &= all(int(pf[i, t].min()) == int(pf[i, t].max()) == (_expected(acts[i], t) + bump) % 256 for i in range(b) for t in range(8))
    worker.serve_request(None)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_sharded_worker_cfg_pairs_world4(tmp_path):
    ctx = mp.get_context("spawn")
    q, port = ctx.Queue(), _free_port()
    ps = [ctx.Process(target=_run_pairs, args=(r, 4, port, str(tmp_path), q)) for r in range(4)]
    [p.start() for p in ps]
    res = [q.get(timeout=180) for _ in ps]
    [p.join(60) for p in ps]
    assert all(ok for _, ok in res), res
