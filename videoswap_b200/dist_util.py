"""Multi-GPU plumbing for the denoising path: one process per GPU, independent editing jobs (videos / prompts) are
partitioned over ranks with no data-path collective; timing is the max over ranks (torch.distributed, NCCL on GPUs,
gloo in the CPU tests)."""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_jobs(n_jobs: int, rank: int, world: int) -> List[int]:
    """Contiguous, balanced partition of job indices: rank r gets jobs [start_r, start_{r+1})."""
    base, extra = divmod(n_jobs, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_latents(latents: torch.Tensor) -> List[torch.Tensor]:
    """All ranks receive every rank's result latents (used when a caller wants all edited videos on one host)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [latents]
    out = [torch.empty_like(latents) for _ in range(dist.get_world_size())]
    dist.all_gather(out, latents.contiguous())
    return out
