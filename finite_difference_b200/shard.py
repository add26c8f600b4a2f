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


def gather_results(local: Dict[str, "np.ndarray"], owned: np.ndarray, n_total: int, device=None,
                   dst: Optional[int] = None) -> Optional[Dict[str, np.ndarray]]:
    """Gather per-option result columns in book order: onto every rank (``dst=None``, all-gather) or onto rank
    ``dst`` only (the other ranks return ``None`` and skip the unpacking -- the final result gather of a sweep).

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
    # shard sizes differ by at most one option: pad to ceil(n_total / world) without a size exchange
    width = -(-n_total // world) if n_total else 0
    if len(owned) > width:
        raise ValueError("shards must come from shard.partition (balanced to within one option)")
    # one collective per dtype: columns of equal dtype are packed side by side, the option index rides along
    groups: Dict[str, list] = {}
    for k, v in local.items():
        v = np.asarray(v)
        groups.setdefault(v.dtype.str, []).append((k, v.reshape(len(owned), -1)))
    idx_col = np.full((width, 1), -1, dtype=np.int64)
    idx_col[: len(owned), 0] = owned
    groups.setdefault(np.dtype(np.int64).str, []).append(("__idx__", idx_col[: len(owned)]))
    gathered = {}
    for dt, cols in groups.items():
        packed = np.zeros((width, sum(c.shape[1] for _, c in cols)), dtype=np.dtype(dt))
        if dt == np.dtype(np.int64).str:
            packed[:, -1] = -1                          # padding rows carry index -1
        off = 0
        for _, c in cols:
            packed[: len(owned), off: off + c.shape[1]] = c
            off += c.shape[1]
        src = torch.from_numpy(packed).to(device)
        if dst is None:
            buf = torch.empty((world * width, packed.shape[1]), dtype=src.dtype, device=device)
            dist.all_gather_into_tensor(buf, src)
            gathered[dt] = buf.cpu().numpy()
        else:
            me = dist.get_rank()
            parts = [torch.empty_like(src) for _ in range(world)] if me == dst else None
            dist.gather(src, parts, dst=dst)
            if me == dst:
                gathered[dt] = torch.cat(parts, dim=0).cpu().numpy()
    if dst is not None and dist.get_rank() != dst:
        return None
    idx_all = gathered[np.dtype(np.int64).str][:, -1]
    valid = idx_all >= 0
    rows = idx_all[valid]
    out = {}
    for dt, cols in groups.items():
        # one row scatter per dtype group (not one per column); the columns are returned as views of it
        src = gathered[dt] if valid.all() else gathered[dt][valid]
        full_group = np.empty((n_total, src.shape[1]), dtype=src.dtype)
        full_group[rows] = src
        off = 0
        for k, c in cols:
            w = c.shape[1]
            if k != "__idx__":
                v = np.asarray(local[k])
                out[k] = full_group[:, off: off + w].reshape((n_total,) + v.shape[1:])
            off += w
    return out
