"""Sharding a book across the GPUs of one box.

PDE solves are independent, so a book shards by OPTION with no communication during the
solve; the only exchange is the final result gather (the reference itself is a
serial ``for _, row in config_df.iterrows()`` loop, run_american_scenarios.py:239).  One
process per GPU (torchrun); ``torch.distributed`` (NCCL on the GPU box, gloo in CPU tests)
carries the gather.

Options are dealt to ranks by cost (``num_space_nodes x total time steps`` of all their solves)
in longest-processing-time order, so every rank gets the same number of node-steps to within
one option.  All solves of one option stay on one rank: Richardson / bump epilogues are local.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import numpy as np

__all__ = ["rank_world", "partition", "my_options", "gather_results", "init_distributed"]


def rank_world() -> tuple:
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def partition(costs: Sequence[float], world: int) -> List[np.ndarray]:
    """Deal options to ``world`` ranks.  Returns one ascending index array per rank.

    Uniform costs -> contiguous-stride round robin (rank r gets r, r+world, ...).  Otherwise
    LPT: options sorted by descending cost, each given to the currently lightest rank.
    """
    costs = np.asarray(costs, dtype=np.float64)
    n = len(costs)
    if world <= 1:
        return [np.arange(n, dtype=np.int64)]
    if n == 0:
        return [np.zeros(0, dtype=np.int64) for _ in range(world)]
    if np.all(costs == costs[0]):
        return [np.arange(r, n, world, dtype=np.int64) for r in range(world)]
    order = np.argsort(-costs, kind="stable")
    # snake deal of the sorted list: within 1 option of LPT for the smooth cost distributions
    # of real books, and O(n log n) instead of a Python heap loop over 1e6 options
    rounds = np.arange(n) // world
    pos = np.arange(n) % world
    owner = np.where(rounds % 2 == 0, pos, world - 1 - pos)
    return [np.sort(order[owner == r]).astype(np.int64) for r in range(world)]


def my_options(costs: Sequence[float], rank: Optional[int] = None, world: Optional[int] = None) -> np.ndarray:
    r, w, _ = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    return partition(costs, world)[rank]


def init_distributed(backend: Optional[str] = None):
    """Initialise ``torch.distributed`` from the torchrun environment; no-op for world size 1."""
    import torch
    import torch.distributed as dist

    rank, world, local = rank_world()
    if world == 1 or dist.is_initialized():
        return rank, world, local
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_results(local: Dict[str, "np.ndarray"], owned: np.ndarray, n_total: int, device=None) -> Dict[str, np.ndarray]:
    """All-gather per-option result columns onto every rank, in book order.

    ``local[name]`` holds this rank's values for options ``owned`` (ascending global indices).
    Shards may differ in length by one option; columns are padded to the longest shard.
    """
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        out = {}
        for k, v in local.items():
            full = np.empty((n_total,) + np.asarray(v).shape[1:], dtype=np.asarray(v).dtype)
            full[owned] = v
            out[k] = full
        return out
    world = dist.get_world_size()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_loc = torch.tensor([len(owned)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_loc) for _ in range(world)]
    dist.all_gather(counts, n_loc)
    counts = [int(c.item()) for c in counts]
    width = max(counts)

    def padded(a, dtype):
        t = torch.zeros((width,) + tuple(a.shape[1:]), dtype=dtype, device=device)
        if len(a):
            t[: len(a)] = torch.as_tensor(np.ascontiguousarray(a), device=device)
        return t

    idx_parts = [torch.zeros(width, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(idx_parts, padded(np.asarray(owned, dtype=np.int64), torch.int64))
    out = {}
    for k, v in local.items():
        v = np.asarray(v)
        tdt = torch.from_numpy(np.zeros(0, dtype=v.dtype)).dtype
        parts = [torch.zeros((width,) + tuple(v.shape[1:]), dtype=tdt, device=device) for _ in range(world)]
        dist.all_gather(parts, padded(v, tdt))
        full = np.empty((n_total,) + v.shape[1:], dtype=v.dtype)
        for r in range(world):
            full[idx_parts[r][: counts[r]].cpu().numpy()] = parts[r][: counts[r]].cpu().numpy()
        out[k] = full
    return out
