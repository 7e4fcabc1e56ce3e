"""Data-parallel host logic: trees never share edges, so any partition of the
trees across GPUs is exact for the forward pass; training adds ONE all-reduce (sum) of the flat
gradient per step.  Ranks get contiguous tree ranges balanced by NODE count (tree sizes are
heavy-tailed), build their own local batch, scale the nll sum by 1/B_global and all-reduce."""
from __future__ import annotations

import numpy as np
import torch


def shard_trees(tree_sizes, world_size: int):
    """Contiguous ranges [lo, hi) of trees per rank with near-equal node counts.
    Deterministic; every tree lands on exactly one rank; ranks may be empty when
    world_size > number of trees."""
    sizes = np.asarray(tree_sizes, dtype=np.int64)
    n = len(sizes)
    cum = np.concatenate([[0], np.cumsum(sizes)])
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world_size):
        target = total * r / world_size
        # first tree boundary whose prefix reaches the target, never moving backwards
        b = int(np.searchsorted(cum, target, side="left"))
        b = min(max(b, bounds[-1]), n)
        bounds.append(b)
    bounds.append(n)
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def node_id_base(tree_sizes, lo: int) -> int:
    """Global id of the first node of tree `lo` (dropout masks are keyed on global node ids,
    so they do not depend on the world size)."""
    return int(np.asarray(tree_sizes, dtype=np.int64)[:lo].sum())


def allreduce_flat_(flat_grad: torch.Tensor, group=None):
    """Sum the flat gradient buffer over the data-parallel group (NCCL on GPUs, gloo in tests)."""
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.all_reduce(flat_grad, op=torch.distributed.ReduceOp.SUM, group=group)
    return flat_grad
