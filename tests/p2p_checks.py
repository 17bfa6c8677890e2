"""Parity helpers for the attention controllers: the native explicit-probability attention, the hook through
AnimateDiffUNet3DModel.forward against the oracle's hook, and the device-native controllers replayed against the fixtures
the reference's own AttentionStore / AttentionRefine / AttentionReplace / SpatialBlender produced (oracle/make_golden_p2p.py)."""
from __future__ import annotations

import os

import torch

from oracle import make_golden_p2p as G
from oracle import unet3d_oracle as O
from tests import unet_checks as U
from videoswap_b200 import ops, p2p

GOLD = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda"


def explicit_attention_check(B=3, N=256, NK=None, C=1280, kv_div=1, seed=170):
    """probs = softmax(q k^T / sqrt(d)) to HBM, then O = P V, vs fp32 torch (d = C / 8)."""
    g = torch.Generator().manual_seed(seed)
    NK = N if NK is None else NK
    q = torch.randn((B, N, C), generator=g).half().to(DEV)
    kv = torch.randn((B // kv_div, NK, 2 * C), generator=g).half().to(DEV)
    k, v = kv[..., :C], kv[..., C:]
    probs = ops.attention_probs(q, k, 8, kv_div=kv_div)
    out = ops.attention_apply_probs(probs, v, 8, kv_div=kv_div)
    torch.cuda.synchronize()
    d = C // 8
    qh = q.float().reshape(B, N, 8, d).transpose(1, 2)
    kh = k.float().repeat_interleave(kv_div, 0).reshape(B, NK, 8, d).transpose(1, 2)
    vh = v.float().repeat_interleave(kv_div, 0).reshape(B, NK, 8, d).transpose(1, 2)
    pr = (qh @ kh.transpose(-1, -2) * d ** -0.5).softmax(-1)
    ref = (pr @ vh).transpose(1, 2).reshape(B, N, C)
    return {"probs_err": (probs.float() - pr).abs().max().item(), "row_sum_err": (probs.float().sum(-1) - 1).abs().max().item(),
            "out_err": (out.float() - ref).abs().max().item(), "out_ref": ref.abs().max().item()}


class _Recorder:
    """Oracle-side twin of p2p.AttentionStore's `forward` (device-agnostic): keeps what the hook saw."""

    def __init__(self, edit=None):
        self.maps, self.edit = [], edit

    def __call__(self, probs, is_cross, place):
        self.maps.append((place, is_cross, probs.detach().clone()))
        return self.edit(probs, is_cross, place) if self.edit else probs


def unet_hook_vs_oracle(Fr=2, hw=16, edit=False):
    """One UNet forward ([2,...] CFG batch) with a controller registered: the maps the native hook delivers and the epsilon
    (after an in-place edit of the conditional half when `edit`) vs the oracle's hook."""
    m, sd = U.get_model()
    x = U.randn((2, 4, Fr, hw, hw), 2).half()
    ehs = U.randn((2, 16, 77, 768), 3).half()

    def edit_fn(probs, is_cross, place):      # a controller that really changes the maps: flatten the cond half of cross maps
        if is_cross:
            h = probs.shape[0] // 2
            probs[h:] = 0.5 * probs[h:] + 0.5 / probs.shape[-1]
        return probs

    class Ctl(p2p.AttentionControl):
        def __init__(self):
            super().__init__()
            self.LOW_RESOURCE = True
            self.seen = []

        def forward(self, attn, is_cross, place):
            self.seen.append((place, is_cross, attn.detach().clone()))
            return edit_fn(attn, is_cross, place) if edit else attn
    ctl = Ctl()
    n = p2p.register_attention_control(m, ctl)
    try:
        out = m(x.cuda(), 981, ehs.cuda(), return_dict=False)[0]
        torch.cuda.synchronize()
    finally:
        p2p.register_attention_control(m, None)
    rec = _Recorder(edit_fn if edit else None)
    O.ATTN_HOOK = rec
    try:
        with torch.no_grad():
            ref = O.unet_forward(sd, O.OracleConfig(), x.float(), 981, ehs.float())
    finally:
        O.ATTN_HOOK = None
    assert len(ctl.seen) == len(rec.maps), (len(ctl.seen), len(rec.maps))
    worst = 1e9
    for (p1, c1, a), (p2_, c2, b) in zip(ctl.seen, rec.maps):
        assert (p1, c1) == (p2_, c2) and tuple(a.shape) == tuple(b.shape), ((p1, c1, a.shape), (p2_, c2, b.shape))
        worst = min(worst, U.psnr(a, b))
    return {"registered": n, "calls": len(ctl.seen), "min_map_psnr": worst, "eps_psnr": U.psnr(out, ref),
            "order": [(p_, c) for p_, c, _ in ctl.seen]}


def replay_vs_reference_fixture(kind="refine"):
    """The scenario of oracle/make_golden_p2p.py replayed on the GPU with the device-native controllers."""
    g = torch.load(os.path.join(GOLD, f"p2p_{kind}.pt"))
    c = g["controller"]
    store = p2p.AttentionStore()
    store.LOW_RESOURCE = True
    for step in range(G.N_STEPS):
        for li, (place, is_cross) in enumerate(G.LAYERS):
            store(G.synth_map(step, li, G.FRAMES, is_cross, 1).half().to(DEV), is_cross, place)
        store.step_callback(G.synth_latents(step, 1).half().to(DEV))

    def blender(alpha, se, choose):
        b = p2p.SpatialBlender(alpha, th=c["th"], NUM_DDIM_STEPS=G.N_STEPS, prompt_choose=choose)
        b.start_blend, b.end_blend = se
        return b
    common = dict(num_steps=G.N_STEPS, cross_replace_alpha=c["cross_replace_alpha"], self_replace_steps=0.0,
                  latent_blend=blender(c["latent_alpha_layers"], c["latent_start_end"], "both"), additional_attention_store=store,
                  attention_blend=blender(c["attn_alpha_layers"], c["attn_start_end"], "source"), image_height=G.IMG, image_width=G.IMG)
    ctl = p2p.AttentionRefine(c["mapper"], c["alphas"], **common) if kind == "refine" else p2p.AttentionReplace(c["mapper"], **common)
    ctl.num_self_replace = tuple(c["num_self_replace"])
    map_err, lat_err, mask_mismatch = 0.0, 0.0, 0
    for step in range(G.N_STEPS):
        for li, (place, is_cross) in enumerate(G.LAYERS):
            attn = G.synth_map(step, li, 2 * G.FRAMES, is_cross, 2).half().to(DEV)
            out = ctl(attn, is_cross, place)
            if li in g["edited"][step]:
                map_err = max(map_err, (out[G.FRAMES:].float().cpu() - g["edited"][step][li]).abs().max().item())
        x = ctl.step_callback(G.synth_latents(step, 2).half().to(DEV))
        lat_err = max(lat_err, (x.float().cpu() - g["latents"][step]).abs().max().item())
    torch.cuda.synchronize()
    masks = [m.float().cpu() for m in ctl.latent_blend.mask_list]
    for a, b in zip(masks, g["latent_masks"]):
        mask_mismatch += int((a != b).sum().item())
    return {"map_err": map_err, "latent_err": lat_err, "mask_mismatch": mask_mismatch, "mask_pixels": sum(m.numel() for m in masks),
            "n_masks": (len(masks), len(g["latent_masks"]))}
