"""Architecture spec of the reference's AnimateDiffUNet3DModel (SD-1.5 UNet + AnimateDiff motion modules):
parameter names/shapes exactly as the reference's `state_dict()` produces them
(videoswap/models/animatediff_models/unet.py:32-255 and unet_blocks.py of the reference), so that
`load_state_dict` / ED-LoRA weight merging (utils/convert_edlora_to_diffusers.py:36-79) keep working.
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass, field, asdict
from typing import Dict, List, Sequence, Tuple


@dataclass
class UNetConfig:
    sample_size: int = 64
    in_channels: int = 4
    out_channels: int = 4
    block_out_channels: Tuple[int, ...] = (320, 640, 1280, 1280)
    layers_per_block: int = 2
    attention_head_dim: int = 8            # == number of heads (reference quirk, unet.py:158)
    cross_attention_dim: int = 768
    norm_num_groups: int = 32
    norm_eps: float = 1e-5
    # AnimateDiff additions (options/model_cfg/inference.yml of the reference)
    use_motion_module: bool = True
    motion_module_resolutions: Tuple[int, ...] = (1, 2, 4, 8)
    motion_module_mid_block: bool = False
    motion_module_decoder_only: bool = False
    motion_num_attention_heads: int = 8
    temporal_position_encoding_max_len: int = 24

    def to_dict(self):
        return asdict(self)

    @property
    def time_embed_dim(self) -> int:
        return self.block_out_channels[0] * 4

    def down_has_motion(self, i: int) -> bool:
        return self.use_motion_module and (2 ** i in self.motion_module_resolutions) and not self.motion_module_decoder_only

    def up_has_motion(self, i: int) -> bool:
        return self.use_motion_module and (2 ** (3 - i) in self.motion_module_resolutions)

    def up_resnet_in_channels(self, i: int, j: int) -> Tuple[int, int]:
        """(channels of the running tensor, channels of the popped skip) for up_blocks.i.resnets.j
        (unet_blocks.py:551-556 of the reference)."""
        boc = list(self.block_out_channels)
        rev = boc[::-1]
        n = len(boc)
        out_c = rev[i]
        prev = rev[max(i - 1, 0)] if i > 0 else rev[0]
        in_c = rev[min(i + 1, n - 1)]
        nl = self.layers_per_block + 1
        skip = in_c if j == nl - 1 else out_c
        run = prev if j == 0 else out_c
        return run, skip


def _resnet(sh, p, cin, cout, temb):
    sh[p + ".norm1.weight"] = (cin,)
    sh[p + ".norm1.bias"] = (cin,)
    sh[p + ".conv1.weight"] = (cout, cin, 3, 3)
    sh[p + ".conv1.bias"] = (cout,)
    sh[p + ".time_emb_proj.weight"] = (cout, temb)
    sh[p + ".time_emb_proj.bias"] = (cout,)
    sh[p + ".norm2.weight"] = (cout,)
    sh[p + ".norm2.bias"] = (cout,)
    sh[p + ".conv2.weight"] = (cout, cout, 3, 3)
    sh[p + ".conv2.bias"] = (cout,)
    if cin != cout:
        sh[p + ".conv_shortcut.weight"] = (cout, cin, 1, 1)
        sh[p + ".conv_shortcut.bias"] = (cout,)


def _attn(sh, p, c, ctx):
    sh[p + ".to_q.weight"] = (c, c)
    sh[p + ".to_k.weight"] = (c, ctx)
    sh[p + ".to_v.weight"] = (c, ctx)
    sh[p + ".to_out.0.weight"] = (c, c)
    sh[p + ".to_out.0.bias"] = (c,)


def _ff(sh, p, c):
    sh[p + ".net.0.proj.weight"] = (8 * c, c)
    sh[p + ".net.0.proj.bias"] = (8 * c,)
    sh[p + ".net.2.weight"] = (c, 4 * c)
    sh[p + ".net.2.bias"] = (c,)


def _norm(sh, p, c):
    sh[p + ".weight"] = (c,)
    sh[p + ".bias"] = (c,)


def _transformer(sh, p, c, ctx):
    _norm(sh, p + ".norm", c)
    sh[p + ".proj_in.weight"] = (c, c, 1, 1)
    sh[p + ".proj_in.bias"] = (c,)
    q = p + ".transformer_blocks.0"
    _attn(sh, q + ".attn1", c, c)
    _norm(sh, q + ".norm1", c)
    _attn(sh, q + ".attn2", c, ctx)
    _norm(sh, q + ".norm2", c)
    _ff(sh, q + ".ff", c)
    _norm(sh, q + ".norm3", c)
    sh[p + ".proj_out.weight"] = (c, c, 1, 1)
    sh[p + ".proj_out.bias"] = (c,)


def _motion(sh, p, c, pe_len):
    p = p + ".temporal_transformer"
    _norm(sh, p + ".norm", c)
    sh[p + ".proj_in.weight"] = (c, c)
    sh[p + ".proj_in.bias"] = (c,)
    q = p + ".transformer_blocks.0"
    for i in (0, 1):
        a = f"{q}.attention_blocks.{i}"
        _attn(sh, a, c, c)
        sh[a + ".processor.pos_encoder.pe"] = (1, pe_len, c)     # buffer (test.py:63 key remap targets this name)
    for i in (0, 1):
        _norm(sh, f"{q}.norms.{i}", c)
    _ff(sh, q + ".ff", c)
    _norm(sh, q + ".ff_norm", c)
    sh[p + ".proj_out.weight"] = (c, c)
    sh[p + ".proj_out.bias"] = (c,)


def unet_param_shapes(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    sh: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    boc = list(cfg.block_out_channels)
    n = len(boc)
    temb = cfg.time_embed_dim
    ctx = cfg.cross_attention_dim
    pe = cfg.temporal_position_encoding_max_len
    sh["conv_in.weight"] = (boc[0], cfg.in_channels, 3, 3)
    sh["conv_in.bias"] = (boc[0],)
    sh["time_embedding.linear_1.weight"] = (temb, boc[0])
    sh["time_embedding.linear_1.bias"] = (temb,)
    sh["time_embedding.linear_2.weight"] = (temb, temb)
    sh["time_embedding.linear_2.bias"] = (temb,)
    cout = boc[0]
    for i in range(n):
        cin, cout = cout, boc[i]
        p = f"down_blocks.{i}"
        cross = i < n - 1
        for j in range(cfg.layers_per_block):
            if cross:
                _transformer(sh, f"{p}.attentions.{j}", cout, ctx)
        for j in range(cfg.layers_per_block):
            _resnet(sh, f"{p}.resnets.{j}", cin if j == 0 else cout, cout, temb)
        for j in range(cfg.layers_per_block):
            if cfg.down_has_motion(i):
                _motion(sh, f"{p}.motion_modules.{j}", cout, pe)
        if i < n - 1:
            sh[f"{p}.downsamplers.0.conv.weight"] = (cout, cout, 3, 3)
            sh[f"{p}.downsamplers.0.conv.bias"] = (cout,)
    c = boc[-1]
    _transformer(sh, "mid_block.attentions.0", c, ctx)
    _resnet(sh, "mid_block.resnets.0", c, c, temb)
    _resnet(sh, "mid_block.resnets.1", c, c, temb)
    if cfg.use_motion_module and cfg.motion_module_mid_block:
        _motion(sh, "mid_block.motion_modules.0", c, pe)
    rev = boc[::-1]
    for i in range(n):
        p = f"up_blocks.{i}"
        cross = i > 0
        out_c = rev[i]
        for j in range(cfg.layers_per_block + 1):
            if cross:
                _transformer(sh, f"{p}.attentions.{j}", out_c, ctx)
        for j in range(cfg.layers_per_block + 1):
            run, skip = cfg.up_resnet_in_channels(i, j)
            _resnet(sh, f"{p}.resnets.{j}", run + skip, out_c, temb)
        for j in range(cfg.layers_per_block + 1):
            if cfg.up_has_motion(i):
                _motion(sh, f"{p}.motion_modules.{j}", out_c, pe)
        if i < n - 1:
            sh[f"{p}.upsamplers.0.conv.weight"] = (out_c, out_c, 3, 3)
            sh[f"{p}.upsamplers.0.conv.bias"] = (out_c,)
    _norm(sh, "conv_norm_out", boc[0])
    sh["conv_out.weight"] = (cfg.out_channels, boc[0], 3, 3)
    sh["conv_out.bias"] = (cfg.out_channels,)
    return sh


def adapter_param_shapes(embedding_channels=1280, channels=(320, 640, 1280, 1280), mid_dim=128):
    """SparsePointAdapter state_dict (videoswap/models/adapter_model.py:50-70 of the reference)."""
    sh = OrderedDict()
    for l, ch in enumerate(channels):
        sh[f"model_list.{l}.mlp.0.weight"] = (mid_dim, embedding_channels)
        sh[f"model_list.{l}.mlp.0.bias"] = (mid_dim,)
        sh[f"model_list.{l}.mlp.2.weight"] = (ch, mid_dim)
        sh[f"model_list.{l}.mlp.2.bias"] = (ch,)
    return sh
