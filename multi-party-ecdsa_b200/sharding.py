"""Sharding of independent work units (operands, proofs, signing sessions) across the GPUs of
one box, and the single gather of fixed-size result records that ends a batch.

Units share nothing at run time, so the partition is a contiguous block split with no data-path
collective; the only exchange is one all-gather of per-rank records (NCCL over NVLink on the
GPU box, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple


def unit_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of `total` units owned by `rank` (sizes differ by at most 1)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def session_range(total_sessions: int, rank: int, world: int, parties: int = 2) -> Tuple[int, int]:
    """Unit range when units come in sessions of `parties` co-resident units (both signers of a
    GG20 session live on the same GPU so their messages never leave device memory)."""
    lo, hi = unit_range(total_sessions, rank, world)
    return lo * parties, hi * parties


def gather_records(records, world: int):
    """All-gather equal-shaped per-rank record tensors -> (world, ...) tensor on every rank."""
    import torch
    import torch.distributed as dist
    if world == 1 or not dist.is_initialized():
        return records.unsqueeze(0)
    parts = [torch.empty_like(records) for _ in range(world)]
    dist.all_gather(parts, records.contiguous())
    return torch.stack(parts)
