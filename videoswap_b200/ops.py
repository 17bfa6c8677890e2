"""Thin torch-tensor wrappers over the per-kernel C-ABI entry points (torch is only used for device memory and the
current stream).  All activations are NHWC fp16 CUDA tensors.  Used by the parity tests and the host-side model."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib

GEGLU_GRANULE = 128
EPI_LINEAR, EPI_GEGLU = 0, 1


def set_option(name: str, value: int):
    """Runtime options of the library, e.g. set_option("attn_tc", 0) forces the mma.sync attention kernel."""
    _lib.call("vs_set_option", name.encode(), int(value))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk16(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float16 and t.is_contiguous(), "expected contiguous fp16 CUDA tensor"


def _chk32(*ts):
    for t in ts:
        if t is not None:
            assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), "expected contiguous fp32 CUDA tensor"


def gemm(A, W, bias=None, residual=None, A2=None, rowvec=None, pix_per_batch=1, mode=EPI_LINEAR, force_bn=0):
    """out[M, N] = [A | A2] @ W^T (+bias +rowvec[row // pix_per_batch] +residual); GEGLU mode expects packed W/bias."""
    _chk16(A, W, residual, A2)
    _chk32(bias, rowvec)
    M, K1 = A.shape
    K2 = 0 if A2 is None else A2.shape[1]
    N = W.shape[0]
    assert W.shape[1] == K1 + K2
    out = torch.empty((M, N // 2 if mode == EPI_GEGLU else N), dtype=torch.float16, device=A.device)
    _lib.call("vs_gemm", _stream(), _p(A), K1, _p(A2), K2, _p(W), M, N, _p(bias), _p(rowvec), pix_per_batch,
              _p(residual), _p(out), mode, force_bn)
    return out


def pack_conv3x3(w):
    """[Co, Ci, 3, 3] fp16 -> [Co, 9*Ci] tap-major."""
    _chk16(w)
    co, ci = w.shape[:2]
    out = torch.empty((co, 9 * ci), dtype=torch.float16, device=w.device)
    _lib.call("vs_pack_conv3x3", _stream(), _p(w), co, ci, _p(out))
    return out


def pack_geglu(w, b):
    """FeedForward.net.0.proj [8C, C] / [8C] -> value/gate interleaved in GEGLU_GRANULE-row (128) granules (+ fp32 bias)."""
    _chk16(w, b)
    hidden, K = w.shape[0] // 2, w.shape[1]
    wout = torch.empty_like(w)
    bout = torch.empty((2 * hidden,), dtype=torch.float32, device=w.device)
    _lib.call("vs_pack_geglu", _stream(), _p(w), _p(b), hidden, K, _p(wout), _p(bout))
    return wout, bout


def conv3x3(x, w_packed, bias=None, x2=None, rowvec=None, imgs_per_batch=1, residual=None):
    """x: [N, H, W, C1] (+x2 [N, H, W, C2] channel concat), w_packed [Co, 9*(C1+C2)] -> [N, H, W, Co]."""
    _chk16(x, w_packed, x2, residual)
    _chk32(bias, rowvec)
    n, H, W, C1 = x.shape
    C2 = 0 if x2 is None else x2.shape[3]
    co = w_packed.shape[0]
    out = torch.empty((n, H, W, co), dtype=torch.float16, device=x.device)
    _lib.call("vs_conv3x3", _stream(), _p(x), C1, _p(x2), C2, _p(w_packed), n, H, W, co, _p(bias), _p(rowvec),
              imgs_per_batch, _p(residual), _p(out))
    return out


def conv3x3_s2(x, w_packed, bias=None):
    _chk16(x, w_packed)
    n, H, W, Ci = x.shape
    co = w_packed.shape[0]
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    scratch = torch.empty((n * Ho * Wo, 9 * Ci), dtype=torch.float16, device=x.device)
    out = torch.empty((n, Ho, Wo, co), dtype=torch.float16, device=x.device)
    _lib.call("vs_conv3x3_s2", _stream(), _p(x), n, H, W, Ci, _p(w_packed), co, _p(bias), _p(scratch), _p(out))
    return out


def upsample_conv3x3(x, w, bias=None):
    """nearest-2x + conv3x3 (pad 1) as four sub-pixel convs: x [N, H, W, C], w [Co, C, 3, 3] (unpacked) -> [N, 2H, 2W, Co]."""
    _chk16(x, w)
    _chk32(bias)
    n, H, W, Ci = x.shape
    co = w.shape[0]
    wsub = torch.empty((16 * co * Ci,), dtype=torch.float16, device=x.device)
    out = torch.empty((n, 2 * H, 2 * W, co), dtype=torch.float16, device=x.device)
    _lib.call("vs_upsample_conv3x3", _stream(), _p(x), n, H, W, Ci, _p(w), co, _p(bias), _p(wsub), _p(out))
    return out


def groupnorm(x1, gamma, beta, groups, eps, imgs_per_set=1, silu=False, x2=None):
    """x1 [N, H, W, C1] (+ x2) -> normalised [N, H, W, C1+C2]; statistics over imgs_per_set images x (C/groups)."""
    _chk16(x1, x2)
    _chk32(gamma, beta)
    n, H, W, c1 = x1.shape
    c2 = 0 if x2 is None else x2.shape[3]
    sums = torch.empty((n // imgs_per_set, groups, 2), dtype=torch.float32, device=x1.device)
    out = torch.empty((n, H, W, c1 + c2), dtype=torch.float16, device=x1.device)
    _lib.call("vs_groupnorm", _stream(), _p(x1), c1, _p(x2), c2, n, H * W, imgs_per_set, groups, eps, _p(gamma), _p(beta),
              int(silu), _p(sums), _p(out))
    return out


def layernorm(x, gamma, beta, pe=None, hw=1, F=1):
    _chk16(x)
    _chk32(gamma, beta, pe)
    rows, Cc = x.shape
    out = torch.empty_like(x)
    _lib.call("vs_layernorm", _stream(), _p(x), rows, Cc, _p(gamma), _p(beta), _p(pe), hw, F, _p(out))
    return out


def ln_linear(x, W, gamma, beta, bias=None, pe=None, hw=1, frames=1, mode=EPI_LINEAR):
    """LayerNorm(x)(+pe[(row // hw) % frames]) @ W^T + bias with the norm folded into the GEMM (C in 320/640/1280)."""
    _chk16(x, W)
    _chk32(gamma, beta, bias, pe)
    M, Cc = x.shape
    N = W.shape[0]
    dev = x.device
    wf = torch.empty_like(W)
    u = torch.empty((N,), dtype=torch.float32, device=dev)
    c = torch.empty((N,), dtype=torch.float32, device=dev)
    cpe = torch.empty((pe.shape[0], N), dtype=torch.float32, device=dev) if pe is not None else None
    stats = torch.empty((M, 2), dtype=torch.float32, device=dev)
    out = torch.empty((M, N // 2 if mode == EPI_GEGLU else N), dtype=torch.float16, device=dev)
    _lib.call("vs_ln_linear", _stream(), _p(x), M, Cc, _p(W), _p(bias), N, _p(gamma), _p(beta), _p(pe),
              0 if pe is None else pe.shape[0], hw, frames, mode, _p(wf), _p(u), _p(c), _p(cpe), _p(stats), _p(out))
    return out


def linear_ln_linear(x0, W0, b0, W, gamma, beta, residual=None, bias=None, mode=EPI_LINEAR):
    """x = x0 @ W0^T + b0 (+ residual) [fp16]; out = LayerNorm(x) @ W^T + bias (or GEGLU), the LayerNorm's row statistics
    coming from the first GEMM's epilogue.  Returns (x, out)."""
    _chk16(x0, W0, W, residual)
    _chk32(b0, gamma, beta, bias)
    M, K0 = x0.shape
    Cc, N = W0.shape[0], W.shape[0]
    dev = x0.device
    x = torch.empty((M, Cc), dtype=torch.float16, device=dev)
    wf = torch.empty_like(W)
    u = torch.empty((N,), dtype=torch.float32, device=dev)
    c = torch.empty((N,), dtype=torch.float32, device=dev)
    cap = 16
    parts = torch.empty((cap, M, 2), dtype=torch.float32, device=dev)
    out = torch.empty((M, N // 2 if mode == EPI_GEGLU else N), dtype=torch.float16, device=dev)
    _lib.call("vs_linear_ln_linear", _stream(), _p(x0), M, K0, _p(W0), _p(b0), _p(residual), Cc, _p(x), _p(W), _p(bias), N,
              _p(gamma), _p(beta), mode, _p(wf), _p(u), _p(c), _p(parts), cap, _p(out))
    return x, out


def attention(q, k, v, heads, kv_div=1):
    """q [B, Nq, h*d], k/v [Bk, Nk, h*d] (may be strided views with contiguous last dim) -> [B, Nq, h*d]."""
    B, nq, Cc = q.shape
    nk = k.shape[1]
    d = Cc // heads
    for t in (q, k, v):
        assert t.dtype == torch.float16 and t.stride(2) == 1
    out = torch.empty((B, nq, Cc), dtype=torch.float16, device=q.device)
    _lib.call("vs_attention", _stream(), _p(q), q.stride(1), _p(k), k.stride(1), _p(v), v.stride(1), _p(out), Cc, B, nq, nk,
              heads, d, q.stride(0), k.stride(0), nq * Cc, kv_div)
    return out


def attention_probs(q, k, heads, kv_div=1):
    """softmax(q k^T / sqrt(d)) as a tensor [B, heads, Nq, Nk] fp16 (explicit-probability path of the attention controllers)."""
    B, nq, Cc = q.shape
    nk = k.shape[1]
    probs = torch.empty((B, heads, nq, nk), dtype=torch.float16, device=q.device)
    _lib.call("vs_attention_probs", _stream(), _p(q), q.stride(1), _p(k), k.stride(1), _p(probs), B, nq, nk, heads, Cc // heads,
              q.stride(0), k.stride(0), kv_div)
    return probs


def attention_apply_probs(probs, v, heads, kv_div=1):
    """O = P V: probs [B, heads, Nq, Nk] fp16, v [Bk, Nk, heads*d] -> [B, Nq, heads*d]."""
    _chk16(probs)
    B, _, nq, nk = probs.shape
    Cc = v.shape[2]
    out = torch.empty((B, nq, Cc), dtype=torch.float16, device=probs.device)
    _lib.call("vs_attention_apply_probs", _stream(), _p(probs), _p(v), v.stride(1), _p(out), Cc, B, nq, nk, heads, Cc // heads,
              v.stride(0), nq * Cc, kv_div)
    return out


def temporal_attention(qkv, heads):
    """qkv [B, F, HW, 3C] -> [B, F, HW, C]: attention over the F axis for every (b, pixel, head)."""
    _chk16(qkv)
    B, F, HW, C3 = qkv.shape
    out = torch.empty((B, F, HW, C3 // 3), dtype=torch.float16, device=qkv.device)
    _lib.call("vs_temporal_attention", _stream(), _p(qkv), _p(out), B, F, HW, C3 // 3, heads)
    return out


def conv_in(x, w, bias):
    _chk16(x, w)
    _chk32(bias)
    n, H, W, ci = x.shape
    co = w.shape[0]
    out = torch.empty((n, H, W, co), dtype=torch.float16, device=x.device)
    _lib.call("vs_conv_in", _stream(), _p(x), n, H, W, ci, _p(w), _p(bias), co, _p(out))
    return out


def upsample2x(x):
    _chk16(x)
    n, H, W, Cc = x.shape
    out = torch.empty((n, 2 * H, 2 * W, Cc), dtype=torch.float16, device=x.device)
    _lib.call("vs_upsample2x", _stream(), _p(x), n, H, W, Cc, _p(out))
    return out


def ddim_coefficients(alpha_t: float, alpha_prev: float):
    """(c_x, c_e) with x_prev = c_x * x + c_e * eps  (DDIM, eta = 0)."""
    import math
    c_x = math.sqrt(alpha_prev) / math.sqrt(alpha_t)
    c_e = math.sqrt(1.0 - alpha_prev) - math.sqrt(alpha_prev) * math.sqrt(1.0 - alpha_t) / math.sqrt(alpha_t)
    return c_x, c_e


def cfg_ddim_step(eps2, latents, guidance, alpha_t=None, alpha_prev=None, cfg=True, out=None, coef=None):
    """coef: optional device tensor [2] fp32 (c_x, c_e) instead of host alphas (CUDA-graph replayable)."""
    assert eps2.dtype == latents.dtype and eps2.is_contiguous() and latents.is_contiguous()
    is_f32 = int(latents.dtype == torch.float32)
    if out is None:
        out = torch.empty_like(latents)
    if coef is not None:
        _chk32(coef)
        _lib.call("vs_cfg_ddim_step_dev", _stream(), _p(eps2), _p(latents), is_f32, latents.numel(), int(cfg), float(guidance),
                  _p(coef), _p(out))
        return out
    _lib.call("vs_cfg_ddim_step", _stream(), _p(eps2), _p(latents), is_f32, latents.numel(), int(cfg), float(guidance),
              float(alpha_t), float(alpha_prev), _p(out))
    return out


def adapter_level(w0, b0, w1, b1, point_embedding, tracks, h, w, rate, point_mask=None, coord_fp16=True, scale=1.0):
    """One level of SparsePointAdapter: returns the NHWC fp16 map [F, h, w, C]."""
    _chk16(w0, b0, w1, b1)
    _chk32(point_embedding, tracks)
    mid, E = w0.shape
    Cc = w1.shape[0]
    F, P = tracks.shape[:2]
    ws = torch.empty((mid + Cc + P * (mid + Cc),), dtype=torch.float32, device=w0.device)
    out = torch.empty((F, h, w, Cc), dtype=torch.float16, device=w0.device)
    _lib.call("vs_adapter_level", _stream(), _p(w0), _p(b0), _p(w1), _p(b1), E, mid, Cc, _p(point_embedding), _p(tracks),
              _p(point_mask), F, P, h, w, float(rate), int(coord_fp16), float(scale), _p(ws), _p(out))
    return out
