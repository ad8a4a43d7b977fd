"""1x128 activation quantiser, mirroring ``sglang_per_token_group_quant_fp8``
(python/sglang/srt/layers/dense/gemms/fp8/fp8_kernel.py:386-463 -> sgl_per_token_group_quant_fp8,
3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-520)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib

fp8_dtype = torch.float8_e4m3fn
fp8_max = 448.0
fp8_min = -448.0


def create_per_token_group_quant_fp8_output_scale(x_shape, device, group_size: int, column_major_scales: bool,
                                                  scale_tma_aligned: bool) -> torch.Tensor:
    """fp8_kernel.py:395-428: scale buffer; column-major + TMA-aligned means storage [k/group][m padded to 4] viewed
    as [m, k/group]."""
    m = 1
    for s in x_shape[:-1]:
        m *= s
    groups = x_shape[-1] // group_size
    if column_major_scales:
        if scale_tma_aligned:
            aligned = (m + 3) // 4 * 4
            return torch.empty((groups, aligned), device=device, dtype=torch.float32).permute(-1, -2)[:m, :]
        return torch.empty((groups, m), device=device, dtype=torch.float32).permute(-1, -2)
    return torch.empty(tuple(x_shape[:-1]) + (groups,), device=device, dtype=torch.float32)


def sglang_per_token_group_quant_fp8(x: torch.Tensor, group_size: int, eps: float = 1e-10,
                                     column_major_scales: bool = False,
                                     scale_tma_aligned: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """x bf16 [..., k] contiguous -> (x_q e4m3 [..., k], x_s f32 [m, k/128]).  Only the layout the DeepGEMM path uses
    (group_size 128, column-major TMA-aligned scales; deep_geem.py:86-93) is implemented natively."""
    assert x.shape[-1] % group_size == 0, "the last dimension of `x` cannot be divisible by `group_size`"
    assert x.is_contiguous(), "`x` is not contiguous"
    if not x.is_cuda:
        raise RuntimeError("sglang_per_token_group_quant_fp8: CUDA tensor required (no CPU fallback)")
    if group_size != 128 or x.dtype != torch.bfloat16:
        raise RuntimeError("sglang_per_token_group_quant_fp8: only bf16 input with group_size=128 is supported")
    k = x.shape[-1]
    m = x.numel() // k
    x_q = torch.empty_like(x, dtype=fp8_dtype)
    if column_major_scales:
        x_s = create_per_token_group_quant_fp8_output_scale(x.shape, x.device, group_size, True, scale_tma_aligned)
        stride = x_s.stride(1) if x_s.shape[1] > 1 else ((m + 3) // 4 * 4 if scale_tma_aligned else m)
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().b200_act_quant_1x128(x.data_ptr(), m, k, x_q.data_ptr(), x_s.data_ptr(), stride,
                                                       float(eps), _lib.current_stream()), "per_token_group_quant_fp8")
        return x_q, x_s
    # row-major scales: quantise into a column-major scratch, then transpose (not on the decode path)
    tmp = torch.empty((k // group_size, m), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200_act_quant_1x128(x.data_ptr(), m, k, x_q.data_ptr(), tmp.data_ptr(), m, float(eps),
                                                   _lib.current_stream()), "per_token_group_quant_fp8")
    return x_q, tmp.t().contiguous().view(tuple(x.shape[:-1]) + (k // group_size,))


def per_token_group_quant_fp8_strided(x: torch.Tensor, eps: float = 1e-10) -> Tuple[torch.Tensor, torch.Tensor]:
    """The DeepGEMM-layout quantiser (group 128, column-major TMA-aligned scales) on a producer's output read in place:
    x bf16 ``[m, k]`` with any row stride (a column slice of a wider projection output) or ``[m, heads, 128]`` with any
    row / head strides (a per-head slice).  Returns (x_q e4m3 ``[m, k]`` contiguous, x_s f32 ``[m, k/128]``) -- the same
    bits ``sglang_per_token_group_quant_fp8(x.contiguous(), 128, column_major_scales=True, scale_tma_aligned=True)``
    gives, without the copy."""
    if not x.is_cuda:
        raise RuntimeError("per_token_group_quant_fp8_strided: CUDA tensor required (no CPU fallback)")
    if x.dtype != torch.bfloat16 or x.dim() not in (2, 3) or x.stride(-1) != 1:
        raise RuntimeError("per_token_group_quant_fp8_strided: bf16 [m, k] or [m, heads, 128] with a contiguous last "
                           "dimension expected")
    m = x.shape[0]
    if x.dim() == 3:
        if x.shape[2] != 128:
            raise RuntimeError("per_token_group_quant_fp8_strided: 3-D input must be [m, heads, 128]")
        k = x.shape[1] * 128
        group_stride = x.stride(1) if x.shape[1] > 1 else 128
        row_stride = x.stride(0) if m > 1 else x.shape[1] * group_stride
    else:
        k = x.shape[1]
        assert k % 128 == 0, "the last dimension of `x` cannot be divisible by `group_size`"
        group_stride, row_stride = 128, (x.stride(0) if m > 1 else k)
    x_q = torch.empty((m, k), dtype=fp8_dtype, device=x.device)
    x_s = create_per_token_group_quant_fp8_output_scale((m, k), x.device, 128, True, True)
    stride = x_s.stride(1) if x_s.shape[1] > 1 else (m + 3) // 4 * 4
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().b200_act_quant_1x128_strided(x.data_ptr(), row_stride, group_stride, m, k, x_q.data_ptr(),
                                                           x_s.data_ptr(), stride, float(eps), _lib.current_stream()),
                   "per_token_group_quant_fp8_strided")
    return x_q, x_s
