"""Deterministic, construction-order-independent synthetic weights.  No pretrained checkpoints exist offline
(SURVEY.md 0.5), so parity and benchmarks use the real architecture with weights drawn per key from a generator
seeded by (seed, crc32(key)).  Motion-module `proj_out` is given NON-zero weights on purpose: the reference
zero-initialises it (motion_module.py:76-77), which would make every motion module an identity and leave the
temporal path untested."""
from __future__ import annotations

import math
import zlib
from collections import OrderedDict
from typing import Dict, Tuple

import torch


def temporal_pe_table(length: int, dim: int) -> torch.Tensor:
    """Closed-form sinusoid of the reference's PositionalEncoding (motion_module.py:242-251)."""
    pos = torch.arange(length, dtype=torch.float32)[:, None]
    div = torch.exp(torch.arange(0, dim, 2, dtype=torch.float32) * (-math.log(10000.0) / dim))
    pe = torch.zeros(1, length, dim)
    pe[0, :, 0::2] = torch.sin(pos * div)
    pe[0, :, 1::2] = torch.cos(pos * div)
    return pe


def seeded_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    sd = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith(".pe"):
            sd[name] = temporal_pe_table(shape[1], shape[2])
            continue
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63 - 1))
        if len(shape) >= 2:
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            bound = 1.0 / math.sqrt(fan_in)
            t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif name.endswith(".weight"):          # 1-D weights are norm scales
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:
            t = 0.02 * torch.randn(shape, generator=g)
        sd[name] = t
    return sd
