"""Branch sharding for multi-GPU requests.

Planner branches (rows of `b_action` + their observation) are independent rollouts, so a request of b
branches is split round-robin over the R ranks of one node (rank r takes r, r+R, ...), each rank rolls
out its branches with no per-step communication, and ONE all-gather of the decoded uint8 frames
(`[b_local, F, 3, H, W]`, padded to the per-rank maximum) returns everything to every rank; rank 0 answers
the manager.  The reference does the same split at process level (worker_manager.py:448-469, 555-570).
Backend-agnostic: NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List

import torch
import torch.distributed as dist


def branch_indices(b: int, rank: int, world: int) -> List[int]:
    return list(range(rank, b, world))


def max_local(b: int, world: int) -> int:
    return (b + world - 1) // world


def gather_frames(local: torch.Tensor, b: int, group=None) -> torch.Tensor:
    """local uint8 [b_local, ...] (this rank's branches, in branch_indices order) -> uint8 [b, ...] in request order.
    One collective: all_gather_into_tensor of the zero-padded per-rank blocks."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    m = max_local(b, world)
    shape = tuple(local.shape[1:])
    block = torch.zeros((m,) + shape, dtype=local.dtype, device=local.device)
    block[: local.shape[0]] = local
    out = torch.empty((world * m,) + shape, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, block, group=group)
    res = torch.empty((b,) + shape, dtype=local.dtype, device=local.device)
    for r in range(world):
        idx = branch_indices(b, r, world)
        if idx:
            res[idx] = out[r * m: r * m + len(idx)]
    return res


# ----------------------------------------------------------------------------- CFG-split pairs (fewer branches than GPUs)
def pair_layout(b: int, world: int):
    """Branch placement when the R ranks work as R/2 CFG pairs (ranks 2p, 2p+1 run the two rows of ONE branch's guidance pair,
    SURVEY.md section 8f item 3): branch i -> pair i % P.  Returns (pairs, owner) with owner[i] = (leader rank, local position)."""
    pairs = world // 2
    owner = [(2 * (i % pairs), i // pairs) for i in range(b)]
    return pairs, owner


def pair_branch_indices(b: int, rank: int, world: int) -> List[int]:
    pairs = world // 2
    return list(range(rank // 2, b, pairs))


def gather_frames_pairs(local: torch.Tensor, b: int, group=None) -> torch.Tensor:
    """Like gather_frames for the pair layout: every rank contributes its (zero-padded) block in ONE all-gather, the pair
    leaders' blocks are the ones read back.  local uint8 [b_local, ...] in pair_branch_indices order."""
    world = dist.get_world_size(group)
    pairs, owner = pair_layout(b, world)
    m = (b + pairs - 1) // pairs
    shape = tuple(local.shape[1:])
    block = torch.zeros((m,) + shape, dtype=local.dtype, device=local.device)
    block[: local.shape[0]] = local
    out = torch.empty((world * m,) + shape, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, block, group=group)
    res = torch.empty((b,) + shape, dtype=local.dtype, device=local.device)
    for i, (leader, pos) in enumerate(owner):
        res[i] = out[leader * m + pos]
    return res
