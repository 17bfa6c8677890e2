"""TEST INFRASTRUCTURE ONLY.  The frame-sharded / CFG-split decomposition of the denoising step, restated on
top of the CPU oracle with torch.distributed collectives (gloo in the tests): it pins WHAT has to be exchanged --

  * 5-D GroupNorm of ResnetBlock3D / conv_norm_out (resnet.py:166,177; unet.py:474): all-reduce of the per-(batch, group)
    (sum, sum of squares) over the frame shards;
  * motion module (motion_module.py:138-162): per-frame GroupNorm locally, then the module body on ALL frames x this rank's
    1/k of the pixels (frames <-> pixels exchange in, and back out before the residual add);
  * CFG: the two halves run on two rank groups and all-gather their noise predictions before the combine

-- independently of the CUDA implementation (videoswap_b200/csrc/unet.cu + comm.cu), which implements the same exchanges
with NCCL.  Everything else of the UNet is frame-local and is the unmodified oracle code."""
from __future__ import annotations

import contextlib
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import unet3d_oracle as O


def _all_gather(t: torch.Tensor, group) -> List[torch.Tensor]:
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t.contiguous(), group=group)
    return out


@contextlib.contextmanager
def frame_sharded(group, shard: int, k: int):
    """Inside this context O.unet_forward treats `sample` as frames [shard F/k, (shard+1) F/k) of the video."""

    def gn5d(sd, p, x, groups, eps):
        b, c = x.shape[:2]
        xg = x.reshape(b, groups, -1)
        stats = torch.stack([xg.sum(-1), (xg * xg).sum(-1)])              # [2, B, G]
        dist.all_reduce(stats, group=group)                                # <- the exchange
        n = xg.shape[-1] * k
        mean = stats[0] / n
        var = stats[1] / n - mean * mean
        y = (xg - mean[..., None]) * torch.rsqrt(var[..., None] + eps)
        shape = (1, c) + (1,) * (x.dim() - 2)
        return y.reshape(x.shape) * sd[p + ".weight"].reshape(shape) + sd[p + ".bias"].reshape(shape)

    def motion(sd, p, x, cfg):
        p = p + ".temporal_transformer"
        b, c, f_loc, h, w = x.shape
        hw = h * w
        assert b == 1 and hw % k == 0
        xb = x.permute(0, 2, 1, 3, 4).reshape(f_loc, c, h, w)
        t = O._gn_frame(sd, p + ".norm", xb, 32, 1e-6)                    # per frame: local
        t = t.permute(0, 2, 3, 1).reshape(f_loc, hw, c)
        full = torch.cat(_all_gather(t, group))                            # <- frames -> pixels ([F, hw, C], keep my pixels)
        f = f_loc * k
        px = slice(shard * hw // k, (shard + 1) * hw // k)
        t = full[:, px]                                                    # [F, hw/k, C]
        t = O._lin(sd, p + ".proj_in", t)
        q = p + ".transformer_blocks.0"
        pe = O.temporal_pe(f, c).to(x)
        for i in (0, 1):
            n = O._ln(sd, f"{q}.norms.{i}", t)
            n = n.permute(1, 0, 2) + pe[None]                              # [(pixels), F, C]
            a = O._attn(sd, f"{q}.attention_blocks.{i}", n, n, cfg.motion_heads)
            t = t + a.permute(1, 0, 2)
        t = t + O._geglu_ff(sd, q + ".ff", O._ln(sd, q + ".ff_norm", t))
        t = O._lin(sd, p + ".proj_out", t)                                 # [F, hw/k, C]
        back = torch.cat(_all_gather(t, group), dim=1)                     # <- pixels -> frames ([F, hw, C], keep my frames)
        mine = back[shard * f_loc:(shard + 1) * f_loc]
        out = mine.reshape(f_loc, h, w, c).permute(0, 3, 1, 2) + xb
        return out.reshape(1, f_loc, c, h, w).permute(0, 2, 1, 3, 4)

    saved = O._gn5d, O.motion_module
    O._gn5d, O.motion_module = gn5d, motion
    try:
        yield
    finally:
        O._gn5d, O.motion_module = saved


def denoise_step_sharded(sd, cfg: O.OracleConfig, latents_full: torch.Tensor, t: int, n_steps: int, ehs2: torch.Tensor,
                         guidance: float, residuals_full: Optional[List[torch.Tensor]], plan, frame_group, cfg_group):
    """One CFG step of ONE video on plan.world ranks; returns this rank's frames of the new latents.  Mirrors
    VideoSwapPipeline.step_sharded (residuals: un-duplicated per-frame maps [(F),C,h,w])."""
    frames = latents_full.shape[2]
    r = plan.frame_range(frames)
    lat = latents_full[:, :, r.start:r.stop]
    res = [m[r.start:r.stop] for m in residuals_full] if residuals_full is not None else None
    e = ehs2[plan.cfg_index:plan.cfg_index + 1] if plan.cfg_ranks == 2 else ehs2
    x = lat if plan.cfg_ranks == 2 else torch.cat([lat] * 2)
    if plan.frame_shards > 1:
        with frame_sharded(frame_group, plan.frame_shard, plan.frame_shards):
            eps = O.unet_forward(sd, cfg, x, t, e, res)
    else:
        eps = O.unet_forward(sd, cfg, x, t, e, [torch.cat([m] * x.shape[0]) for m in res] if res is not None else None)
    if plan.cfg_ranks == 2:
        eps = torch.cat(_all_gather(eps, cfg_group))                       # <- uncond first
    return O.DDIM().step(O.cfg_combine(eps, guidance), t, lat, n_steps)
