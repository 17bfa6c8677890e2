"""A/B of attention kernel options on one box: interleaved repetitions, medians.
usage: gpu_attn_ab.py [--reps N] cfg [cfg ...]   with cfg = name=v[:name=v...], e.g.  attn_persist=0  attn_persist=1:attn_poly=2
Shapes: the C2 launches (BF = 32): L0 self (88 % of the attention FLOPs), L0 cross (77 keys), L1 self, L1 cross."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from videoswap_b200 import ops  # noqa: E402

args = sys.argv[1:]
reps = 11
if args and args[0] == "--reps":
    reps = int(args[1])
    args = args[2:]
cfgs = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in a.split(":")) for a in args]
names = sorted({k for c in cfgs for k in c})
dev = "cuda"
shapes = {"L0 self 4096x4096 d40": (32, 4096, 4096, 320, 1), "L0 cross 4096x77 d40": (32, 4096, 77, 320, 16),
          "L1 self 1024x1024 d80": (32, 1024, 1024, 640, 1), "L1 cross 1024x77 d80": (32, 1024, 77, 640, 16)}
out = {}
for label, (B, N, NK, C, kvdiv) in shapes.items():
    if NK == N:
        qkv = torch.randn(B, N, 3 * C, device=dev).half()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    else:
        q = torch.randn(B, N, C, device=dev).half()
        kv = torch.randn(B // kvdiv, NK, 2 * C, device=dev).half()
        k, v = kv[..., :C], kv[..., C:]
    times = [[] for _ in cfgs]
    ref = None
    for r in range(reps + 2):
        for ci, c in enumerate(cfgs):
            for n in names:
                ops.set_option(n, c.get(n, {"attn_persist": 1, "attn_poly": 1, "attn_handoff": 1, "attn_tc": 1, "attn_epiwg": 1, "attn_pingpong": 1, "attn_ptmem": 0}.get(n, 0)))
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            o = ops.attention(q, k, v, 8, kv_div=kvdiv)
            b.record()
            torch.cuda.synchronize()
            if r == 0:
                if ref is None:
                    ref = o.float()
                else:
                    d = (o.float() - ref).abs().max().item()
                    assert d < 2e-2, (label, c, d)          # variants agree with each other (parity proper: kernel checks)
            if r >= 2:
                times[ci].append(a.elapsed_time(b) * 1e3)
    flops = 4.0 * B * 8 * N * NK * (C // 8)
    for ci, c in enumerate(cfgs):
        t = sorted(times[ci])[len(times[ci]) // 2]
        out[f"{label} | {args[ci]}"] = {"median_us": t, "min_us": min(times[ci]), "tflops": flops / t / 1e6}
        print(f"{label}: {args[ci]}: median {t:8.1f} us  ({flops / t / 1e6:6.1f} TFLOP/s)  min {min(times[ci]):8.1f}", flush=True)
# ---- cycle breakdown of the instrumented persistent kernel at L0 self (softmax warp 4, lane 0; mean over the CTAs)
import ctypes as C  # noqa: E402

from videoswap_b200 import _lib  # noqa: E402
qkv = torch.randn(32, 4096, 960, device=dev).half()
q, k, v = qkv[..., :320], qkv[..., 320:640], qkv[..., 640:]
for epiwg in (0, 1):
    for n in ("attn_persist", "attn_debug", "attn_epiwg", "attn_poly"):
        ops.set_option(n, {"attn_persist": 2, "attn_debug": 1, "attn_epiwg": epiwg, "attn_poly": 1}[n])
    for _ in range(2):
        ops.attention(q, k, v, 8)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * (148 * 16))()
    _lib.call("vs_debug_read", buf, 148 * 16)
    a = torch.tensor(list(buf), dtype=torch.float64).reshape(148, 16)
    names = ["total", "wait_S", "wait_turn", "wait_Pbuf", "exp_phase", "epilogue", "wait_S_first_tile", "tiles",
             "mma_total", "mma_wait_Q", "mma_wait_KV", "mma_wait_P0", "mma_wait_P1", "mma_wait_Vones", "mma_wait_Obuf", "qk_wait_Sbuf"]
    mean, mx, mn = a.mean(0), a.max(0).values, a.min(0).values
    out[f"debug_epiwg{epiwg}"] = {nm: {"mean": float(mean[i]), "min": float(mn[i]), "max": float(mx[i])} for i, nm in enumerate(names)}
    print(f"debug epiwg={epiwg}: " + "  ".join(f"{nm}={mean[i]:.0f}[{mn[i]:.0f}..{mx[i]:.0f}]" for i, nm in enumerate(names)), flush=True)
ops.set_option("attn_debug", 0)
ops.set_option("attn_epiwg", 1)
ops.set_option("attn_persist", 1)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "attn_ab.json"), "w") as f:
    json.dump(out, f, indent=1)
