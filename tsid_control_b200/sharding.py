"""Env sharding over the GPUs of one box and the one collective of the tick.

Every env is an independent QP, so the batch is partitioned by contiguous env index, N/G per rank,
with no exchange on the solve.  The only collective is an all-gather of the per-tick
diagnostics (status, iteration count): int32 [N_local, 2] per rank, 8 bytes per env.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of rank's envs; the first n_total % world ranks get one more."""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_diagnostics(status: torch.Tensor, iters: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
    """All-gather (status, iters) of every rank's shard -> int32 [N_total, 2] in global env order.
    Shards must have equal length (pad the last one) — true for the weak-scaling layout bench.py uses."""
    diag = torch.stack([status.to(torch.int32), iters.to(torch.int32)], dim=1).contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return diag
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * diag.shape[0], 2), dtype=torch.int32, device=diag.device)
    dist.all_gather_into_tensor(out, diag)
    return out


def status_histogram(diag: torch.Tensor) -> torch.Tensor:
    """Counts of HQP status -1..4 over the gathered diagnostics (6 bins)."""
    return torch.bincount((diag[:, 0] + 1).clamp(0, 5).to(torch.int64), minlength=6)
