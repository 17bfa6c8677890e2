"""Per-shape microbenchmarks of the hot kernels at the C2 workload (BF = 32, 64x64 latent): CUDA-event timing,
L2 flushed between reps.  Writes gpurun_out/shape_bench.json."""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from videoswap_b200 import ops  # noqa: E402

DEV = "cuda"
FLUSH = None


def timeit(fn, reps=10, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush:
            FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = {}
    NI = 32
    levels = [(320, 4096), (640, 1024), (1280, 256), (1280, 64)]
    # ---- GEMMs: (name, M, N, K, mode, count/step)
    gemms = []
    for li, (C, hw) in enumerate(levels):
        M = NI * hw
        ntr = [5, 5, 5, 1][li]
        nmo = 5
        gemms += [(f"L{li}_qkv", M, 3 * C, C, 0, ntr + 2 * nmo), (f"L{li}_proj", M, C, C, 0, 5 * ntr + 4 * nmo),
                  (f"L{li}_ff1_geglu", M, 8 * C, C, 1, ntr + nmo), (f"L{li}_ff2", M, C, 4 * C, 0, ntr + nmo)]
    for name, M, N, K, mode, cnt in gemms:
        A = torch.randn(M, K, device=DEV).half()
        W = (torch.randn(N, K, device=DEV) / math.sqrt(K)).half()
        b = torch.randn(N, device=DEV)
        if mode == 1:
            W, b = ops.pack_geglu(W, b.half())
        R = torch.randn(M, N, device=DEV).half() if mode == 0 else None
        ms = timeit(lambda: ops.gemm(A, W, bias=b, residual=R, mode=mode))
        res[name] = {"ms": ms, "tflops": 2.0 * M * N * K / ms / 1e9, "count": cnt, "ms_total": ms * cnt, "shape": [M, N, K]}
        print(name, res[name], flush=True)
        # A/B: CTA pairs (cta_group::2) forced on / off
        ops.set_option("gemm_pair", 2)
        ms2 = timeit(lambda: ops.gemm(A, W, bias=b, residual=R, mode=mode))
        ops.set_option("gemm_pair", 0)
        ms3 = timeit(lambda: ops.gemm(A, W, bias=b, residual=R, mode=mode))
        ops.set_option("gemm_pair", 1)
        res[f"{name}_pair"] = {"ms_pair": ms2, "ms_single": ms3, "tflops_pair": 2.0 * M * N * K / ms2 / 1e9}
        print(f"{name}_pair", res[f"{name}_pair"], flush=True)
    # ---- convs
    for name, n, H, ci, co, cnt in [("conv_L0_320", NI, 64, 320, 320, 8), ("conv_L0_960_320", NI, 64, 960, 320, 1),
                                    ("conv_L0_640_320", NI, 64, 640, 320, 2), ("conv_L1_640", NI, 32, 640, 640, 8),
                                    ("conv_L1_1920_640", NI, 32, 1920, 640, 1), ("conv_L2_1280", NI, 16, 1280, 1280, 9),
                                    ("conv_L2_2560_1280", NI, 16, 2560, 1280, 2), ("conv_L3_1280", NI, 8, 1280, 1280, 8),
                                    ("conv_L3_2560_1280", NI, 8, 2560, 1280, 3)]:
        x = torch.randn(n, H, H, ci, device=DEV).half()
        w = ops.pack_conv3x3((torch.randn(co, ci, 3, 3, device=DEV) / math.sqrt(9 * ci)).half())
        b = torch.randn(co, device=DEV)
        ms = timeit(lambda: ops.conv3x3(x, w, bias=b))
        ops.set_option("gemm_pair", 0)
        ms2 = timeit(lambda: ops.conv3x3(x, w, bias=b))
        ops.set_option("gemm_pair", 1)
        fl = 2.0 * n * H * H * co * 9 * ci
        res[name] = {"ms": ms, "tflops": fl / ms / 1e9, "count": cnt, "ms_total": ms * cnt, "ms_single_cta": ms2,
                     "tflops_single_cta": fl / ms2 / 1e9}
        print(name, res[name], flush=True)
    # ---- attention
    for li, (C, hw) in enumerate(levels):
        qkv = torch.randn(NI, hw, 3 * C, device=DEV).half()
        q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        ms = timeit(lambda: ops.attention(q, k, v, 8))
        fl = 4.0 * NI * 8 * hw * hw * (C // 8)
        res[f"attn_L{li}"] = {"ms": ms, "tflops": fl / ms / 1e9, "count": [5, 5, 5, 1][li], "ms_total": ms * [5, 5, 5, 1][li]}
        print(f"attn_L{li}", res[f"attn_L{li}"], flush=True)
        if li < 2:
            ops.set_option("attn_tc", 0)
            ms2 = timeit(lambda: ops.attention(q, k, v, 8))
            ops.set_option("attn_tc", 1)
            res[f"attn_L{li}_mma"] = {"ms": ms2, "tflops": fl / ms2 / 1e9}
            print(f"attn_L{li}_mma", res[f"attn_L{li}_mma"], flush=True)
        kv = torch.randn(2, 77, 2 * C, device=DEV).half()
        qq = torch.randn(NI, hw, C, device=DEV).half()
        ms = timeit(lambda: ops.attention(qq, kv[..., :C], kv[..., C:], 8, kv_div=16))
        res[f"xattn_L{li}"] = {"ms": ms, "tflops": 4.0 * NI * 8 * hw * 77 * (C // 8) / ms / 1e9, "count": [5, 5, 5, 1][li]}
        print(f"xattn_L{li}", res[f"xattn_L{li}"], flush=True)
        t = torch.randn(2, 16, hw, 3 * C, device=DEV).half()
        ms = timeit(lambda: ops.temporal_attention(t, 8))
        res[f"tattn_L{li}"] = {"ms": ms, "gbs": 8.0 * NI * hw * C / ms / 1e6, "count": 10}
        print(f"tattn_L{li}", res[f"tattn_L{li}"], flush=True)
    # ---- norms
    for li, (C, hw) in enumerate(levels):
        x = torch.randn(NI * hw, C, device=DEV).half()
        g = torch.ones(C, device=DEV)
        ms = timeit(lambda: ops.layernorm(x, g, g))
        res[f"ln_L{li}"] = {"ms": ms, "gbs": 4.0 * NI * hw * C / ms / 1e6}
        print(f"ln_L{li}", res[f"ln_L{li}"], flush=True)
        h = int(math.isqrt(hw))
        x4 = x.reshape(NI, h, h, C)
        ms = timeit(lambda: ops.groupnorm(x4, g, g, 32, 1e-5, imgs_per_set=16, silu=True))
        res[f"gn5d_L{li}"] = {"ms": ms, "gbs": 6.0 * NI * hw * C / ms / 1e6}
        print(f"gn5d_L{li}", res[f"gn5d_L{li}"], flush=True)
        ms = timeit(lambda: ops.groupnorm(x4, g, g, 32, 1e-6, imgs_per_set=1, silu=False))
        res[f"gnframe_L{li}"] = {"ms": ms, "gbs": 6.0 * NI * hw * C / ms / 1e6}
        print(f"gnframe_L{li}", res[f"gnframe_L{li}"], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "shape_bench.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
