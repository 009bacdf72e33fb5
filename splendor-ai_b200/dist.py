"""Multi-GPU plumbing: games shard independently by contiguous id range; the only collective is one all-reduce of the
int64[16] counter block (wins per seat, ties, games, steps, endings, score sum) - SURVEY.md section 8e.  RNG is keyed by
GLOBAL game id, so any number of ranks produces identical per-game results."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world():
    """(rank, local_rank, world_size) from the torchrun environment (single process if absent)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend: str | None = None) -> tuple[int, int, int]:
    rank, local_rank, size = world()
    if size > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": torch.device("cuda", local_rank)} if backend == "nccl" else {}
        dist.init_process_group(backend, rank=rank, world_size=size, **kw)
    return rank, local_rank, size


def shutdown():
    if dist.is_initialized():
        dist.destroy_process_group()


def shard_range(total: int, rank: int, size: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of `total` games owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(total, size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_counters(counters: torch.Tensor) -> torch.Tensor:
    """Sum the counter block over ranks (in place); no-op for a single process."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


def allreduce_max(value: float, device) -> float:
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def barrier():
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
