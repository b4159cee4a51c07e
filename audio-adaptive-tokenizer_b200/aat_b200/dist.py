"""Multi-GPU plumbing: utterances are independent, so ranks never exchange data on the path itself
(SURVEY.md §8e).  This module only decides which utterances a rank owns and performs the single
dataset-level allreduce of the ``dim + 1`` column sums.  One process per GPU (``torchrun``)."""
from __future__ import annotations

import os
from typing import List, Sequence

import numpy as np


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_by_duration(n_samples: Sequence[int], world_size: int) -> List[np.ndarray]:
    """Greedy longest-first partition of utterance indices so that every rank streams (almost) the
    same number of samples.  Deterministic; indices inside a shard stay in ascending order."""
    n = np.asarray(n_samples, dtype=np.int64)
    order = np.argsort(-n, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    shards: List[list] = [[] for _ in range(world_size)]
    for i in order.tolist():
        r = int(np.argmin(load))
        shards[r].append(i)
        load[r] += n[i]
    return [np.asarray(sorted(s), dtype=np.int64) for s in shards]


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous equal split of ``n_items`` identical work units (weak-scaling benchmark shards)."""
    lo = (n_items * rank) // world_size
    hi = (n_items * (rank + 1)) // world_size
    return lo, hi


def allreduce_sum_(tensor, group=None):
    """In-place SUM allreduce when a process group exists; identity otherwise."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=group)
    return tensor
