"""Oracle: block-scaled W8A8 FP8 GEMM and its 1x128 activation quantiser (TEST INFRASTRUCTURE ONLY).

numpy restatements; reference paths relative to /root/reference.
  * w8a8_block_fp8_matmul  -- python/sglang/test/test_block_fp8.py:89-141 (native_w8a8_block_fp8_matmul), which is the
    accumulation contract of 3rdparty/deep_gemm/deep_gemm/include/deep_gemm/fp8_gemm.cuh:278-331 (fp32 inner sum per
    128-deep k-block, fp32 outer sum weighted by sA * sB).
  * per_token_group_quant_fp8 -- python/sglang/test/test_block_fp8.py:15-40 (native_per_token_group_quant_fp8) ==
    3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-437 (absmax = max(eps, |x|), s = absmax / 448,
    q = clamp(x / s)).
Pinned by tests/golden/block_fp8_*.npz generated from the reference's own Python (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

from .fp8 import FP8_MAX, f32_to_e4m3


def w8a8_block_fp8_matmul(a_f32, a_scales, b_f32, b_scales, block_n: int = 128, block_k: int = 128,
                          round_bf16: bool = False):
    """a_f32 [m,k] (decoded e4m3 values), a_scales [m, k/128], b_f32 [n,k], b_scales [ceil(n/128), k/128] -> f32 [m,n]."""
    a = np.asarray(a_f32, dtype=np.float32)
    b = np.asarray(b_f32, dtype=np.float32)
    a_s = np.asarray(a_scales, dtype=np.float32)
    b_s = np.asarray(b_scales, dtype=np.float32)
    m, k = a.shape
    n = b.shape[0]
    k_tiles = (k + block_k - 1) // block_k
    n_tiles = (n + block_n - 1) // block_n
    assert a_s.shape == (m, k_tiles) and b_s.shape == (n_tiles, k_tiles)
    c = np.zeros((m, n), dtype=np.float32)
    for i in range(k_tiles):
        ks = slice(i * block_k, min((i + 1) * block_k, k))
        for j in range(n_tiles):
            ns = slice(j * block_n, min((j + 1) * block_n, n))
            part = (a[:, ks] @ b[ns, ks].T).astype(np.float32)
            s = (a_s[:, i:i + 1] * b_s[j, i]).astype(np.float32)
            c[:, ns] = (c[:, ns] + part * s).astype(np.float32)
    if round_bf16:
        from .fp8 import round_to_bf16
        return round_to_bf16(c)
    return c


def per_token_group_quant_fp8(x_f32, group_size: int = 128, eps: float = 1e-10):
    """x [m,k] f32 (bf16-representable values) -> (codes u8 [m,k], scales f32 [m, k/group])."""
    x = np.asarray(x_f32, dtype=np.float32)
    m, k = x.shape
    assert k % group_size == 0
    xg = x.reshape(m, k // group_size, group_size)
    amax = np.maximum(np.max(np.abs(xg), axis=-1), np.float32(eps)).astype(np.float32)
    s = (amax / FP8_MAX).astype(np.float32)
    q = np.clip((xg / s[..., None]).astype(np.float32), -FP8_MAX, FP8_MAX)
    return f32_to_e4m3(q).reshape(m, k), s


def fused_qk_rmsnorm(qkv, q_dim: int, kv_dim: int, q_weight, kv_weight, eps: float):
    """f-2: FusedRMSNorm.forward (python/sglang/srt/layers/layernorm.py:276-303 -> flashinfer norm.cuh:474-630):
    out = bf16(float(x) * rsqrt(mean(x^2) + eps) * float(w)) for the q_a columns [0, q_dim) and the kv_a columns
    [q_dim, q_dim + kv_dim) of every row, evaluated in float64 and rounded once (the kernel's fp32 evaluation may differ
    by one bf16 ulp).  Returns (q_a normalised [m, q_dim], kv_a normalised [m, kv_dim]) as float32 holding bf16 values."""
    from .fp8 import round_to_bf16
    x = np.asarray(qkv, dtype=np.float64)
    outs = []
    for lo, hi, w in ((0, q_dim, q_weight), (q_dim, q_dim + kv_dim, kv_weight)):
        part = x[:, lo:hi]
        rcp = 1.0 / np.sqrt((part * part).mean(axis=1, keepdims=True) + eps)
        outs.append(round_to_bf16((part * rcp * np.asarray(w, dtype=np.float64)[None, :]).astype(np.float32)))
    return outs[0], outs[1]
