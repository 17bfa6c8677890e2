"""Host-side mirror of the reference's `SparsePointAdapter` (videoswap/models/adapter_model.py:50-136) and of the
denoising part of `VideoSwapPipeline` (videoswap/pipelines/pipeline_videoswap.py:427-619 `__call__`, :622-721 `invert`).

Scope: the loop body -- CFG batch duplication, UNet forward, CFG combine, scheduler step, adapter
residual window -- runs on the native kernels.  Text encoding (CLIP), VAE encode/decode, prompt/LoRA handling and the
attention controllers are callers on either side of this path (8f) and are NOT part of this package: the pipeline
takes `prompt_embeds` / `latents` tensors and returns latents.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import torch
from torch import nn

from . import ops
from .scheduler import DDIMInverseScheduler, DDIMScheduler
from .spec import adapter_param_shapes
from .unet import AnimateDiffUNet3DModel, _Holder
from .weights import seeded_state_dict


@dataclass
class TuneAVideoPipelineOutput:
    videos: torch.Tensor


@dataclass
class TuneAVideoInversionPipelineOutput:
    latents: torch.Tensor


class _MLP(nn.Module):
    def __init__(self, in_dim, out_dim, mid_dim):
        super().__init__()
        self.mlp = nn.ModuleList([_Holder((mid_dim, in_dim)), nn.SiLU(), _Holder((out_dim, mid_dim))])


class SparsePointAdapter(nn.Module):
    """Same constructor/state_dict as the reference; `forward` returns NHWC fp16 maps (one per level) produced by the
    native MLP + splat kernels.  `as_nchw=True` returns the reference's [(F), C, h, w] layout instead."""

    def __init__(self, embedding_channels=1280, channels=(320, 640, 1280, 1280), downsample_rate=(8, 16, 32, 64),
                 mid_dim=128, init: str = "seeded"):
        super().__init__()
        self.model_list = nn.ModuleList([_MLP(embedding_channels, ch, mid_dim) for ch in channels])
        self.downsample_rate = list(downsample_rate)
        self.channels = list(channels)
        self.radius = 2
        if init == "seeded":
            self.load_state_dict(seeded_state_dict(adapter_param_shapes(embedding_channels, channels, mid_dim), seed=5))
        self.eval()

    @torch.no_grad()
    def forward(self, point_tracker, size, point_embedding, index_list=None, drop_rate=0.0, loss_type="global",
                scale: float = 1.0, coord_fp16: bool = True, as_nchw: bool = False) -> List[torch.Tensor]:
        if self.training:
            raise NotImplementedError("adapter training (loss mask / point dropout, adapter_model.py:72-95,107-109) is a "
                                      "'next' row (SURVEY 8f-3)")
        dev = point_embedding.device
        if dev.type != "cuda":
            raise RuntimeError("SparsePointAdapter (videoswap_b200) runs on CUDA only")
        tracks = point_tracker.squeeze(0) if point_tracker.dim() == 4 else point_tracker
        emb = point_embedding.squeeze(0) if point_embedding.dim() == 3 else point_embedding
        w, h = size
        nf, npts = tracks.shape[:2]
        mask = None
        if index_list is not None:
            mask = torch.zeros(npts, dtype=torch.int32, device=dev)
            mask[list(index_list)] = 1
        # the reference casts coordinates to the latents dtype (fp16 at inference, pipeline_videoswap.py:533)
        tr = tracks.to(device=dev, dtype=torch.float32).contiguous()
        if coord_fp16:
            tr = tracks.to(device=dev, dtype=torch.float16).float().contiguous()
        emb32 = emb.to(device=dev, dtype=torch.float32).contiguous()
        out = []
        for lv, mlp in enumerate(self.model_list):
            rate = self.downsample_rate[lv]
            p = [t.detach().to(device=dev, dtype=torch.float16).contiguous()
                 for t in (mlp.mlp[0].weight, mlp.mlp[0].bias, mlp.mlp[2].weight, mlp.mlp[2].bias)]
            m = ops.adapter_level(p[0], p[1], p[2], p[3], emb32, tr, h // rate, w // rate, rate, mask, coord_fp16, scale)
            out.append(m.permute(0, 3, 1, 2).contiguous() if as_nchw else m)
        return out


def _combine(eps, latents, guidance_scale, a_t, a_p, cfg):
    """CFG combine + DDIM update: the fused CUDA kernel (no CPU path)."""
    if not eps.is_cuda:
        raise RuntimeError("the CFG + DDIM update runs on CUDA only")
    return ops.cfg_ddim_step(eps, latents, guidance_scale, a_t, a_p, cfg=cfg)


class VideoSwapPipeline:
    """The denoising loop of the reference pipeline on the native path.  BASELINE.json's `TuneAVideoPipeline` alias."""

    def __init__(self, unet: AnimateDiffUNet3DModel, scheduler: Optional[DDIMScheduler] = None,
                 adapter: Optional[SparsePointAdapter] = None, inverse_scheduler: Optional[DDIMInverseScheduler] = None):
        self.unet = unet
        self.scheduler = scheduler or DDIMScheduler()
        self.inverse_scheduler = inverse_scheduler or DDIMInverseScheduler()
        self.adapter = adapter

    @property
    def device(self):
        return self.unet.device

    def _residuals_for_cfg(self, adapter_state, cfg: bool):
        """NHWC maps [(F),h,w,C] -> the NCHW [(B F),C,h,w] list the UNet surface takes (B = 2 under CFG)."""
        res = []
        for m in adapter_state:
            r = m.permute(0, 3, 1, 2)
            res.append(torch.cat([r, r], dim=0).contiguous() if cfg else r.contiguous())
        return res

    @torch.no_grad()
    def step_sharded(self, latents: torch.Tensor, t, embeds: torch.Tensor, guidance_scale: float, plan,
                     residuals: Optional[List[torch.Tensor]] = None, coef: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The loop body on ONE video split over ranks (dist_util.ShardPlan).  `latents` [1,4,F/k,h,w] and
        `residuals` 4 x [(F/k),C,h,w] hold THIS rank's frames; `embeds` is the full [2,...] (uncond first).  The rank runs the
        UNet on its CFG half (batch 1; GroupNorm statistics and the motion modules exchange inside the library), the two
        halves swap their noise predictions, and both compute the same DDIM update for their frames."""
        from . import dist_util
        cfg = guidance_scale > 1.0
        if cfg and plan.cfg_ranks != 2:
            raise ValueError("frame sharding under CFG needs the CFG split (one batch element per rank)")
        e = embeds[plan.cfg_index:plan.cfg_index + 1] if cfg else embeds
        res = [r for r in residuals] if residuals is not None else None
        eps = self.unet(latents, t, encoder_hidden_states=e, down_block_additional_residuals=res, return_dict=False)[0]
        if cfg:
            eps = dist_util.all_gather_cfg(plan, eps)
        if coef is not None:
            return ops.cfg_ddim_step(eps, latents, guidance_scale, cfg=cfg, coef=coef)
        a_t, a_p = self.scheduler.alphas(t)
        return ops.cfg_ddim_step(eps, latents, guidance_scale, a_t, a_p, cfg=cfg)

    @torch.no_grad()
    def step(self, latents: torch.Tensor, t: int, embeds: torch.Tensor, guidance_scale: float = 7.5,
             residuals: Optional[List[torch.Tensor]] = None) -> torch.Tensor:
        """One loop body (pipeline_videoswap.py:556-587): CFG batch duplication -> UNet -> CFG combine -> DDIM step.
        `embeds` is [2,...] (uncond first) when guidance_scale > 1 else [1,...]; `scheduler.set_timesteps` must have
        been called.  Returns the new latents."""
        cfg = guidance_scale > 1.0
        a_t, a_p = self.scheduler.alphas(t)
        x_in = torch.cat([latents] * 2) if cfg else latents
        x_in = self.scheduler.scale_model_input(x_in, t)
        eps = self.unet(x_in, t, encoder_hidden_states=embeds, down_block_additional_residuals=residuals, return_dict=False)[0]
        return _combine(eps, latents, guidance_scale, a_t, a_p, cfg)

    @torch.no_grad()
    def __call__(self, prompt_embeds: torch.Tensor, latents: torch.Tensor, negative_prompt_embeds: Optional[torch.Tensor] = None,
                 conditions: Optional[Dict] = None, num_inference_steps: int = 50, guidance_scale: float = 7.5,
                 t2i_guidance_scale: float = 1.0, t2i_start: float = 0.0, t2i_end: float = 1.0, controller=None,
                 output_type: str = "latent", return_dict: bool = True, callback=None, callback_steps: int = 1,
                 max_iters: Optional[int] = None):
        """prompt_embeds: [1,77,D] / ED-LoRA [1,16,77,D] (conditional); negative_prompt_embeds same shape (uncond).
        latents [1,4,F,h,w] (e.g. DDIM-inverted).  Mirrors pipeline_videoswap.py:552-601."""
        if output_type != "latent":
            raise NotImplementedError("VAE decode is outside the hot path (SURVEY 8f-4); use output_type='latent'")
        cfg = guidance_scale > 1.0
        dev = latents.device
        if cfg:
            if negative_prompt_embeds is None:
                raise ValueError("classifier-free guidance needs negative_prompt_embeds")
            embeds = torch.cat([negative_prompt_embeds, prompt_embeds], dim=0)     # uncond FIRST (edlora_util.py:190-195)
        else:
            embeds = prompt_embeds
        self.scheduler.set_timesteps(num_inference_steps)
        timesteps = self.scheduler.timesteps
        adapter_state = None
        if conditions is not None:
            if self.adapter is None:
                raise ValueError("conditions given but the pipeline has no adapter")
            # the reference casts tracks AND the point embedding to the latents dtype first (pipeline_videoswap.py:528-533)
            emb = conditions["point_embedding"].to(dev)
            if latents.dtype == torch.float16:
                emb = emb.half()
            adapter_state = self.adapter(conditions["pred_tracks"].to(dev), conditions["img_size"], emb,
                                         index_list=conditions["index_list"], scale=t2i_guidance_scale,
                                         coord_fp16=latents.dtype == torch.float16)
            adapter_state = self._residuals_for_cfg(adapter_state, cfg)
        latents = latents.contiguous()
        for i, t in enumerate(timesteps):
            if max_iters is not None and i >= max_iters:      # truncated schedules (tests / benchmarks)
                break
            res = None
            if adapter_state is not None and len(timesteps) * t2i_start <= i <= len(timesteps) * t2i_end:
                res = list(adapter_state)        # the UNet pops from this list (no clone needed: it never writes to them)
            latents = self.step(latents, t, embeds, guidance_scale, res)
            if controller is not None:           # edit the latents using the attention maps (pipeline_videoswap.py:589-593)
                latents = controller.step_callback(latents).to(latents.dtype)
            if callback is not None and i % callback_steps == 0:
                callback(i, t, latents)
        # the reference always rearranges 'b c f h w -> (b f) c h w' before returning (pipeline_videoswap.py:603-610)
        b, c, f, h, w = latents.shape
        video = latents.permute(0, 2, 1, 3, 4).reshape(b * f, c, h, w)
        if not return_dict:
            return video
        return TuneAVideoPipelineOutput(videos=video)

    @torch.no_grad()
    def invert(self, prompt_embeds: torch.Tensor, latents: torch.Tensor, num_inference_steps: int = 50,
               return_dict: bool = True, controller=None, max_iters: Optional[int] = None):
        """DDIM inversion loop (pipeline_videoswap.py:677-703), guidance_scale = 1 (no CFG).  The UNet is evaluated at
        the inverse scheduler's timestep; which noise levels the step connects is the scheduler's `convention`."""
        self.inverse_scheduler.set_timesteps(num_inference_steps)
        latents = latents.contiguous()
        for i, t in enumerate(self.inverse_scheduler.timesteps):
            if max_iters is not None and i >= max_iters:
                break
            eps = self.unet(latents, t, encoder_hidden_states=prompt_embeds, return_dict=False)[0]
            a_cur, a_next = self.inverse_scheduler.alphas(t)
            latents = ops.cfg_ddim_step(eps, latents, 1.0, a_cur, a_next, cfg=False)
            if controller is not None:           # store the maps / latents of this inversion step (pipeline_videoswap.py:698-702)
                latents = controller.step_callback(latents).to(latents.dtype)
        if not return_dict:
            return latents
        return TuneAVideoInversionPipelineOutput(latents=latents.detach().clone())


class GraphedStep:
    """One loop body (CFG duplication -> UNet -> CFG combine -> DDIM update) captured once in a CUDA graph and replayed
    for every timestep: the timestep and the two DDIM coefficients live in device memory (3 floats uploaded from a
    pinned host buffer before each replay), everything else (weights, embeddings, adapter residuals, workspace) is
    static.  Removes ~750 kernel-launch gaps per step."""

    def __init__(self, pipe: VideoSwapPipeline, latents: torch.Tensor, embeds: torch.Tensor, guidance_scale: float = 7.5,
                 residuals: Optional[List[torch.Tensor]] = None, inverse: bool = False, plan=None):
        """inverse=True: the DDIM-inversion loop body (pipeline_videoswap.py:677-696, no CFG) -- `__call__` then takes the
        inverse scheduler's timesteps and coefficients."""
        self.pipe, self.guidance, self.residuals, self.inverse = pipe, guidance_scale, residuals, inverse
        self.plan = plan if (plan is not None and plan.world > 1) else None    # sharded: latents / residuals are this rank's frames
        if inverse:
            assert guidance_scale <= 1.0 and residuals is None, "the inversion loop runs without CFG and without adapter residuals"
            if pipe.inverse_scheduler.num_inference_steps is None:
                pipe.inverse_scheduler.set_timesteps(pipe.scheduler.num_inference_steps or 50)
        dev = latents.device
        self.lat = latents.clone().contiguous()
        self.embeds = embeds.contiguous()
        self._h = torch.zeros(3, dtype=torch.float32).pin_memory()      # timestep, c_x, c_e
        self._d = torch.zeros(3, dtype=torch.float32, device=dev)
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            self._d.copy_(torch.tensor([1.0, 1.0, 0.0]))
            for _ in range(2):                                          # warm-up: workspace + weights in place
                self._body()
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._body()
        # the graph holds raw pointers into the UNet's workspace arena: forbid its re-allocation while this object lives
        from . import _lib
        self._pinned = pipe.unet._handle
        _lib.call("vs_unet_pin_workspace", self._pinned, 1)

    def __del__(self):
        try:
            if getattr(self, "_pinned", None) is not None:
                from . import _lib
                _lib.lib().vs_unet_pin_workspace(self._pinned, 0)
        except Exception:  # noqa: BLE001
            pass

    def _body(self):
        if self.plan is not None:      # the NCCL exchanges are captured with the kernels
            return self.pipe.step_sharded(self.lat, self._d[0:1], self.embeds, self.guidance, self.plan, self.residuals,
                                          coef=self._d[1:3])
        cfg = self.guidance > 1.0
        x_in = torch.cat([self.lat] * 2) if cfg else self.lat
        res = list(self.residuals) if self.residuals is not None else None
        eps = self.pipe.unet(x_in, self._d[0:1], encoder_hidden_states=self.embeds, down_block_additional_residuals=res,
                             return_dict=False)[0]
        return ops.cfg_ddim_step(eps, self.lat, self.guidance, cfg=cfg, coef=self._d[1:3])

    def __call__(self, latents: torch.Tensor, t: int) -> torch.Tensor:
        """Runs the step at timestep t.  The returned tensor is overwritten by the next call."""
        a_t, a_p = (self.pipe.inverse_scheduler if self.inverse else self.pipe.scheduler).alphas(t)
        c_x, c_e = ops.ddim_coefficients(a_t, a_p)      # the same update serves both directions: x' = c_x x + c_e eps
        # a fresh pinned staging tensor per call: torch's caching host allocator keeps it alive until the async copy ran
        h = torch.tensor([float(t), c_x, c_e], dtype=torch.float32).pin_memory()
        self._d.copy_(h, non_blocking=True)
        if latents.data_ptr() != self.lat.data_ptr():
            self.lat.copy_(latents, non_blocking=True)
        self.graph.replay()
        return self.out


TuneAVideoPipeline = VideoSwapPipeline
