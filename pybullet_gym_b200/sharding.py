"""Multi-GPU plumbing: one process per GPU, each owning an independent slice of the environments.

There is no collective on the step path: envs never interact.  The only exchanges are
(a) the reduction of a few episode statistics and (b) the max-over-ranks of the timed region in
bench.py, both over ``torch.distributed`` (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import Dict, Tuple

import torch
import torch.distributed as dist

STAT_KEYS = ("return_sum", "length_sum", "episodes", "truncated", "nonfinite", "steps", "contact_overflow")


def rank_world() -> Tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard(rank: int, world: int, envs_per_gpu: int) -> Tuple[int, int]:
    """(env_offset, num_envs) of this rank: rank g owns global envs [g*E, (g+1)*E) and therefore the RNG
    streams (seed, g*E + i) -- results do not depend on how many GPUs the job runs on."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return rank * envs_per_gpu, envs_per_gpu


def reduce_stats(stats: Dict[str, float], device="cpu") -> Dict[str, float]:
    """Sum the per-rank episode statistics over all ranks (no-op without an initialised process group)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return dict(stats)
    t = torch.tensor([float(stats.get(k, 0)) for k in STAT_KEYS], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: float(v) for k, v in zip(STAT_KEYS, t.tolist())}


def max_over_ranks(x: float, device="cpu") -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
