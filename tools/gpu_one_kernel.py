"""Runs ONE kernel shape a few times (for ncu captures).  usage: gpu_one_kernel.py {qkv|proj|ff1|conv|attn|ln|gn} [reps]"""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from videoswap_b200 import ops  # noqa: E402

which = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
M = 131072
if which in ("qkv", "proj", "ff1", "ff2"):
    N, K, mode = {"qkv": (960, 320, 0), "proj": (320, 320, 0), "ff1": (2560, 320, 1), "ff2": (320, 1280, 0)}[which]
    A = torch.randn(M, K, device=dev).half()
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).half()
    b = torch.randn(N, device=dev)
    R = torch.randn(M, N, device=dev).half() if which in ("proj", "ff2") else None
    if mode == 1:
        W, b = ops.pack_geglu(W, b.half())
    fn = lambda: ops.gemm(A, W, bias=b if which != "qkv" else None, residual=R, mode=mode)  # noqa: E731
elif which == "conv":
    x = torch.randn(32, 64, 64, 320, device=dev).half()
    w = ops.pack_conv3x3((torch.randn(320, 320, 3, 3, device=dev) / 54).half())
    b = torch.randn(320, device=dev)
    fn = lambda: ops.conv3x3(x, w, bias=b)  # noqa: E731
elif which == "attn":
    qkv = torch.randn(32, 4096, 960, device=dev).half()
    fn = lambda: ops.attention(qkv[..., :320], qkv[..., 320:640], qkv[..., 640:], 8)  # noqa: E731
elif which == "ln":
    x = torch.randn(M, 320, device=dev).half()
    g = torch.ones(320, device=dev)
    fn = lambda: ops.layernorm(x, g, g)  # noqa: E731
elif which == "gn":
    x = torch.randn(32, 64, 64, 320, device=dev).half()
    g = torch.ones(320, device=dev)
    fn = lambda: ops.groupnorm(x, g, g, 32, 1e-5, imgs_per_set=16, silu=True)  # noqa: E731
for _ in range(reps):
    fn()
torch.cuda.synchronize()
print("done", which)
