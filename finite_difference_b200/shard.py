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

__all__ = ["rank_world", "partition", "my_options", "gather_results", "init_distributed", "PackedGather",
           "ShardedBookRun"]


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


class PackedGather:
    """The final result gather of a sharded sweep, on the device.

    Every rank holds its shard's results as one byte buffer of ``width`` rows per column (``n_f64`` float64 columns
    followed by ``n_i32`` int32 columns: ``book.DeviceBookRun.packed``).  One collective (``dist.gather`` onto
    ``dst``, or ``all_gather_into_tensor`` with ``dst=None``) moves the buffers; the receiving rank puts the rows into
    book order with one ``index_select`` per dtype ON THE DEVICE and copies the two blocks to pinned host memory
    (two D2H copies).  No host-side packing, no per-row scatter.

    ``owned``: the global option index of every row of this rank's buffer (``shard.partition``, in the order the
    rows are laid out); the owners are exchanged once here.
    """

    def __init__(self, owned: np.ndarray, n_total: int, width: int, n_f64: int = 6, n_i32: int = 3, device=None,
                 dst: Optional[int] = 0):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.dst, self.n_total, self.width, self.n_f64, self.n_i32 = dst, int(n_total), int(width), n_f64, n_i32
        if device is None:
            device = (torch.device("cuda", torch.cuda.current_device())
                      if (dist.is_initialized() and dist.get_backend() == "nccl") else torch.device("cpu"))
        self.device = torch.device(device)
        self.row_bytes = 8 * n_f64 + 4 * n_i32
        owned = np.asarray(owned, dtype=np.int64)
        if len(owned) > width:
            raise ValueError("width must hold the largest shard")
        if width % 2:
            raise ValueError("width must be even (the int32 columns follow the float64 columns in one buffer)")
        idx = torch.full((width,), -1, dtype=torch.int64, device=self.device)
        idx[: len(owned)] = torch.from_numpy(owned).to(self.device)
        if self.world > 1:
            all_idx = torch.empty((self.world * width,), dtype=torch.int64, device=self.device)
            dist.all_gather_into_tensor(all_idx, idx)
        else:
            all_idx = idx
        self.receives = dst is None or self.rank == dst
        self.collective = "none" if self.world == 1 else ("all_gather_into_tensor" if dst is None else f"gather(dst={dst})")
        if self.receives:
            all_idx_h = all_idx.cpu().numpy()
            pos = np.nonzero(all_idx_h >= 0)[0]
            perm = np.empty(self.n_total, dtype=np.int64)
            perm[all_idx_h[pos]] = pos                     # book index -> row of the gathered layout
            self.perm = torch.from_numpy(perm).to(self.device)
            self.buf = torch.empty((self.world, width * self.row_bytes), dtype=torch.uint8, device=self.device)
            pin = self.device.type == "cuda"
            self.host_f64 = torch.empty((n_f64, self.n_total), dtype=torch.float64, pin_memory=pin)
            self.host_i32 = torch.empty((n_i32, self.n_total), dtype=torch.int32, pin_memory=pin)
            self.d2h_bytes = self.host_f64.numel() * 8 + self.host_i32.numel() * 4

    def exchange(self, packed):
        """The collective and the reordering alone: returns ``(f64 [n_f64, n_total], i32 [n_i32, n_total])`` DEVICE
        tensors in book order on the receiving rank(s), ``None`` elsewhere.  Asynchronous on the current stream."""
        torch, dist = self.torch, self.dist
        w = self.width
        if self.world > 1:
            if self.dst is None:
                dist.all_gather_into_tensor(self.buf.view(-1), packed)
            else:
                dist.gather(packed, list(self.buf.unbind(0)) if self.receives else None, dst=self.dst)
        elif self.receives:
            self.buf[0].copy_(packed)
        if not self.receives:
            return None
        nf, ni = self.n_f64, self.n_i32
        f = self.buf[:, : 8 * nf * w].view(torch.float64).view(self.world, nf, w).permute(1, 0, 2).reshape(nf, self.world * w)
        i = self.buf[:, 8 * nf * w:].view(torch.int32).view(self.world, ni, w).permute(1, 0, 2).reshape(ni, self.world * w)
        return torch.index_select(f, 1, self.perm), torch.index_select(i, 1, self.perm)

    def __call__(self, packed, sync: bool = True):
        """``packed``: uint8 tensor of ``width * row_bytes`` bytes on ``self.device``.  Returns ``(f64 [n_f64, n_total],
        i32 [n_i32, n_total])`` numpy views of the pinned host blocks on the receiving rank(s), ``None`` elsewhere."""
        got = self.exchange(packed)
        if got is None:
            return None
        self.host_f64.copy_(got[0], non_blocking=True)
        self.host_i32.copy_(got[1], non_blocking=True)
        if sync and self.device.type == "cuda":
            self.torch.cuda.current_stream(self.device).synchronize()
        return self.host_f64.numpy(), self.host_i32.numpy()


class ShardedBookRun:
    """One rank's part of a sweep sharded by option over the GPUs of one box -- the multi-GPU form of
    ``AmericanBook.run_fused`` / ``BarrierBook.run_fused``.  ``step()`` = H2D of this shard's per-option inputs from
    pinned host memory, fan-out + solves + epilogue on the device (``DeviceBookRun``), one result collective, and on
    the receiving rank the D2H of the whole book's results in book order.  No communication during the solve
    (the reference is the serial loop at run_american_scenarios.py:239).
    """

    def __init__(self, book_shard, owned: np.ndarray, n_total: int, what: str = "scenario", dst: Optional[int] = 0,
                 device=None, **opts):
        from .book import DeviceBookRun
        world = rank_world()[1]
        width = self.shard_width(n_total, world)
        self.run = book_shard.device_run(what, device=device, width=width, **opts)
        self.gather = PackedGather(owned, n_total, width, len(DeviceBookRun.F64_COLS), len(DeviceBookRun.I32_COLS),
                                   device=self.run.device, dst=dst)
        self.n_total = int(n_total)

    @staticmethod
    def shard_width(n_total: int, world: int) -> int:
        """Rows per column of every rank's packed result buffer: the largest shard, rounded up to an even count."""
        w = -(-int(n_total) // max(world, 1))
        return w + (w & 1)

    def step(self, upload: bool = True):
        from .book import DeviceBookRun
        if upload:
            self.run.upload()
        self.run.launch()
        got = self.gather(self.run.packed)
        if got is None:
            return None
        f64, i32 = got
        out = {k: f64[j] for j, k in enumerate(DeviceBookRun.F64_COLS)}
        out.update({k: i32[j] for j, k in enumerate(DeviceBookRun.I32_COLS)})
        return out
