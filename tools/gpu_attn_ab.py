"""A/B of attention kernel options on one box: interleaved repetitions, medians.  usage: gpu_attn_ab.py name=v1,v2,... [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from videoswap_b200 import ops  # noqa: E402

name, vals = sys.argv[1].split("=")
vals = [int(v) for v in vals.split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 15
dev = "cuda"
shapes = {"L0 self 4096x4096 d40": (32, 4096, 320), "L1 self 1024x1024 d80": (32, 1024, 640)}
for label, (B, N, C) in shapes.items():
    qkv = torch.randn(B, N, 3 * C, device=dev).half()
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    times = {x: [] for x in vals}
    for r in range(reps + 2):
        for x in vals:
            ops.set_option(name, x)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.attention(q, k, v, 8)
            b.record()
            torch.cuda.synchronize()
            if r >= 2:
                times[x].append(a.elapsed_time(b) * 1e3)
    flops = 4.0 * B * 8 * N * N * (C // 8)
    for x in vals:
        t = sorted(times[x])[len(times[x]) // 2]
        print(f"{label}: {name}={x}: median {t:8.1f} us  ({flops / t / 1e6:6.1f} TFLOP/s)  min {min(times[x]):8.1f}", flush=True)
