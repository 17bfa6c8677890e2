"""Pipeline-depth experiment: conv / GEMM time vs. number of smem ring stages and tile width (is the mainloop latency-bound?)."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from videoswap_b200 import ops  # noqa: E402
from tools.gpu_shape_bench import timeit  # noqa: E402

dev = "cuda"
x = torch.randn(32, 64, 64, 320, device=dev).half()
w = ops.pack_conv3x3((torch.randn(320, 320, 3, 3, device=dev) / 54).half())
b = torch.randn(320, device=dev)
fl = 2.0 * 32 * 64 * 64 * 320 * 2880
for st in (0, 5, 4, 3, 2):
    ops.set_option("gemm_stages", st)
    ms = timeit(lambda: ops.conv3x3(x, w, bias=b))
    print(f"conv_L0_320 stages={st}: {ms*1e3:.1f} us  {fl/ms/1e9:.0f} TF", flush=True)
ops.set_option("gemm_stages", 0)
A = torch.randn(8192, 5120, device=dev).half()
for N, bn in ((1280, 0), (1280, 128), (1280, 64)):
    W = (torch.randn(N, 5120, device=dev) / 70).half()
    for st in (0, 4, 3, 2):
        ops.set_option("gemm_stages", st)
        ms = timeit(lambda: ops.gemm(A, W, force_bn=bn))
        print(f"gemm 8192x{N}x5120 bn={bn} stages={st}: {ms*1e3:.1f} us  {2.0*8192*N*5120/ms/1e9:.0f} TF", flush=True)
ops.set_option("gemm_stages", 0)
