"""Multi-GPU plumbing of the denoising path (the reference has no inference-time parallelism): one process
per GPU, ONE video split over the ranks along the two axes the UNet shards on:

  * CFG axis (2-way): the unconditional / conditional halves of the UNet batch are independent through the whole UNet
    (GroupNorm statistics are per batch element); the two noise predictions meet in one all-gather before the combine.
  * frame axis (k-way): every conv, spatial transformer and per-frame norm is frame-local; the 45 cross-frame GroupNorms
    all-reduce their (sum, sum of squares) and each motion module runs frames <-> pixels re-sharded (libvideoswap_b200's
    comm.cu; NCCL over NVLink).

world 2 -> CFG split; world 4 -> CFG x 2 frame shards; world 8 -> CFG x 4 frame shards.  Without CFG (DDIM inversion) all
ranks are frame shards.  The exchanges themselves live in the C library (vs_comm_*, vs_unet_set_frame_shard); this module
only builds the rank plan, creates the communicators (the 128-byte NCCL ids travel over torch.distributed) and offers
the small host-side helpers the bench and the tests share."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional

import torch
import torch.distributed as dist


@dataclass
class ShardPlan:
    world: int = 1
    rank: int = 0
    cfg_ranks: int = 1          # 2: this rank computes ONE half of the CFG batch
    frame_shards: int = 1       # k
    cfg_index: int = 0          # 0 = unconditional half, 1 = conditional half (uncond first, edlora_util.py:190-195)
    frame_shard: int = 0
    frame_comm: Optional[C.c_void_p] = field(default=None, repr=False)    # vs_comm* of the k ranks sharing this half
    cfg_comm: Optional[C.c_void_p] = field(default=None, repr=False)      # vs_comm* of the 2 ranks sharing these frames

    @property
    def frame_group(self) -> List[int]:
        return [self.cfg_index * self.frame_shards + s for s in range(self.frame_shards)]

    @property
    def cfg_group(self) -> List[int]:
        return [c * self.frame_shards + self.frame_shard for c in range(self.cfg_ranks)]

    def frame_range(self, frames: int) -> range:
        if frames % self.frame_shards:
            raise ValueError(f"{frames} frames do not split over {self.frame_shards} frame shards")
        n = frames // self.frame_shards
        return range(self.frame_shard * n, (self.frame_shard + 1) * n)

    def shard_frames(self, t: torch.Tensor, dim: int) -> torch.Tensor:
        r = self.frame_range(t.shape[dim])
        return t.narrow(dim, r.start, len(r)).contiguous()

    def shard_frame_major(self, t: torch.Tensor, frames: int) -> torch.Tensor:
        """[(F), ...] per-frame tensors (adapter maps): this rank's frames."""
        return self.shard_frames(t.reshape(frames, *t.shape[1:]), 0)


def make_plan(world: int, rank: int, cfg: bool = True) -> ShardPlan:
    """Rank layout: rank = cfg_index * k + frame_shard (the ranks of one CFG half are contiguous)."""
    if world == 1:
        return ShardPlan()
    if cfg:
        if world % 2:
            raise ValueError("the CFG split needs an even number of ranks")
        k = world // 2
        return ShardPlan(world, rank, 2, k, rank // k, rank % k)
    return ShardPlan(world, rank, 1, world, 0, rank)


def shard_jobs(n_jobs: int, rank: int, world: int) -> List[int]:
    """Contiguous, balanced partition of independent jobs (replica mode: one video / editing prompt per rank)."""
    base, extra = divmod(n_jobs, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def create_comms(plan: ShardPlan) -> ShardPlan:
    """Creates the NCCL communicators of the plan inside libvideoswap_b200 (one per exchange group).  Every rank draws one
    id PER KIND of group (a ncclUniqueId serves exactly one communicator: rank 0 is the first rank of a frame group AND of
    a CFG pair), one all_gather_object spreads them, each group uses the id its first rank drew for that kind."""
    from . import _lib
    if plan.world == 1:
        return plan
    mine = []
    for _ in range(2):
        buf = (C.c_char * 128)()
        _lib.call("vs_comm_unique_id", buf)
        mine.append(bytes(buf.raw))
    ids: List[Optional[List[bytes]]] = [None] * plan.world
    dist.all_gather_object(ids, mine)
    for name, (group, idb) in select_comm_ids(plan, ids).items():
        h = C.c_void_p()
        _lib.call("vs_comm_create", (C.c_char * 128).from_buffer_copy(idb), group.index(plan.rank), len(group), C.byref(h))
        setattr(plan, name, h)
    return plan


def select_comm_ids(plan: ShardPlan, ids):
    """ids[rank] = [id for a frame group, id for a CFG pair] as drawn by `rank`.  Returns {attribute: (group, id)} for the
    groups of more than one rank this rank belongs to: the id its group's first rank drew for that kind."""
    out = {}
    for kind, (name, group) in enumerate((("frame_comm", plan.frame_group), ("cfg_comm", plan.cfg_group))):
        if len(group) > 1:
            out[name] = (group, ids[group[0]][kind])
    return out


def attach(unet, plan: ShardPlan) -> None:
    """Tells the native UNet which frame shard it is (vs_unet_forward then takes the LOCAL frame count)."""
    from . import _lib
    unet._ensure_handle()
    _lib.call("vs_unet_set_frame_shard", unet._handle, plan.frame_comm, plan.frame_shard, plan.frame_shards)


def all_gather_cfg(plan: ShardPlan, eps: torch.Tensor) -> torch.Tensor:
    """[1, ...] noise prediction of this rank's CFG half -> [2, ...] (uncond first) on both ranks of the pair."""
    from . import _lib
    out = torch.empty((2,) + tuple(eps.shape[1:]), dtype=eps.dtype, device=eps.device)
    e = eps.contiguous()
    _lib.call("vs_comm_all_gather", plan.cfg_comm, torch.cuda.current_stream().cuda_stream, e.data_ptr(), out.data_ptr(),
              e.numel() * e.element_size())
    return out


def gather_frames(latents: torch.Tensor, plan: ShardPlan, dim: int = 2) -> torch.Tensor:
    """Final join: every rank receives the full-length latents (torch.distributed all_gather over the frame shards of its
    CFG half; both halves hold identical latents)."""
    if plan.frame_shards == 1:
        return latents
    parts = [torch.empty_like(latents) for _ in range(plan.world)]
    dist.all_gather(parts, latents.contiguous())
    return torch.cat([parts[r] for r in plan.frame_group], dim=dim)
