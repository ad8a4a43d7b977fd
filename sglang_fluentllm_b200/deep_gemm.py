"""Drop-in for the reference's ``deep_gemm`` dense entry point (3rdparty/deep_gemm/deep_gemm/__init__.py,
jit_kernels/gemm.py:145-226): ``gemm_fp8_fp8_bf16_nt(lhs, rhs, out, use_pdl)`` plus the layout helpers its callers use.
Backed by the AOT sm_100a kernel behind ``b200_fp8_gemm_nt`` (no JIT: shapes are runtime arguments).
"""
from __future__ import annotations

from typing import Tuple

import torch

from . import _lib

_num_sms = None
_workspaces = {}   # (device index, stream handle) -> split-K scratch (fp32 partial slabs; no initialisation needed)
_retired = []      # outgrown scratch buffers: a captured CUDA graph may still hold their addresses, so they are never freed


def _workspace(device: torch.device, nbytes: int):
    """Split-K scratch, one buffer per (device, stream): GEMMs running concurrently on several streams (the reference's
    two-batch overlap, tbo_executor + deep_gemm.set_num_sms) never share slabs.  Grown on demand, except while a CUDA
    graph is being captured: the buffer's address is baked into the graph, so it must exist -- large enough -- before
    capture starts; the runner's warm-up runs see to that (cuda_graph_runner.py:423-427; the reference's JIT compile
    happens in the same warm-up).  Asking for more during capture is an error, not a silent change of tiling."""
    if nbytes == 0:
        return None
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("deep_gemm: split-K workspace of %d bytes requested during CUDA-graph capture; run the same "
                               "call once on this stream before capturing (warm-up) so the scratch exists" % nbytes)
        if ws is not None:
            _retired.append(ws)
        ws = torch.empty(max(nbytes, 64 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def ceil_div(x: int, y: int) -> int:
    return (x + y - 1) // y


def set_num_sms(num_sms: int) -> None:  # jit_kernels/utils.py: kept for API compatibility; grids are tile-shaped
    global _num_sms
    _num_sms = int(num_sms)


def get_num_sms() -> int:
    global _num_sms
    if _num_sms is None:
        _num_sms = _lib.lib().b200_sm_count()
    return _num_sms


def get_tma_aligned_size(x: int, element_size: int) -> int:
    """jit_kernels/utils.py:55-75."""
    alignment = 16 // element_size
    return ceil_div(x, alignment) * alignment


def get_col_major_tma_aligned_tensor(x: torch.Tensor) -> torch.Tensor:
    """Same contract as the reference helper of this name (jit_kernels/utils.py:77-107): the [m, k/128] (or batched
    [b, m, k/128]) scale tensor comes back with the same logical shape but column-major storage -- element (i, j) at
    offset j * m_pad + i -- where m_pad is m rounded up to a 16-byte multiple, which is what a TMA row pitch needs.
    A tensor that already has that storage is returned as is."""
    if x.dim() not in (2, 3):
        raise ValueError("expected a 2-D or 3-D scale tensor, got %d-D" % x.dim())
    batched = x.dim() == 3
    xb = x if batched else x[None]
    nb, m, groups = xb.shape
    m_pad = get_tma_aligned_size(m, x.element_size())
    if xb.stride(1) == 1 and xb.stride(2) == m_pad and (nb == 1 or xb.stride(0) == m_pad * groups):
        return x
    storage = x.new_empty((nb, groups, m_pad))
    col_major = storage.permute(0, 2, 1)[:, :m, :]
    col_major.copy_(xb)
    return col_major if batched else col_major[0]


def gemm_fp8_fp8_bf16_nt(lhs: Tuple[torch.Tensor, torch.Tensor], rhs: Tuple[torch.Tensor, torch.Tensor],
                         out: torch.Tensor, use_pdl: bool = False) -> None:
    """out[m, n] (bf16) = lhs[m, k] @ rhs[n, k]^T with 1x128 LHS scales and 128x128 RHS scales, written in place.
    Same checks as jit_kernels/gemm.py:159-177."""
    lhs, lhs_scales = lhs
    rhs, rhs_scales = rhs
    m, k = lhs.shape
    n, k_ = rhs.shape
    m_, n_ = out.shape
    assert n % 64 == 0 and k % 128 == 0
    assert m == m_ and n == n_ and k == k_
    assert n > 0 and k > 0
    assert lhs_scales.shape == (m, (k + 127) // 128)
    assert rhs_scales.shape == ((n + 127) // 128, (k + 127) // 128)
    assert lhs.dtype == torch.float8_e4m3fn and lhs_scales.dtype == torch.float32
    assert rhs.dtype == torch.float8_e4m3fn and rhs_scales.dtype == torch.float32
    assert out.dtype == torch.bfloat16
    assert lhs.is_contiguous() and rhs.is_contiguous() and out.is_contiguous()
    lhs_scales = get_col_major_tma_aligned_tensor(lhs_scales)
    assert rhs_scales.is_contiguous()
    if m == 0:
        return
    if not lhs.is_cuda:
        raise RuntimeError("gemm_fp8_fp8_bf16_nt: tensors must be on a CUDA device (no CPU fallback)")
    with torch.cuda.device(lhs.device):
        L = _lib.lib()
        ws = _workspace(lhs.device, L.b200_fp8_gemm_workspace_bytes(m, n, k))
        _lib.check(
            L.b200_fp8_gemm_nt(lhs.data_ptr(), lhs_scales.data_ptr(),
                               lhs_scales.stride(1) if lhs_scales.shape[1] > 1 else get_tma_aligned_size(m, 4),
                               rhs.data_ptr(), rhs_scales.data_ptr(), out.data_ptr(), m, n, k,
                               ws.data_ptr() if ws is not None else None, ws.numel() if ws is not None else 0,
                               int(bool(use_pdl)), _lib.current_stream()), "gemm_fp8_fp8_bf16_nt")


def get_m_alignment_for_contiguous_layout() -> int:
    """jit_kernels/utils.py: the row alignment of every group in the contiguous layout (one group per 128-row tile)."""
    return 128


def m_grouped_gemm_fp8_fp8_bf16_nt_contiguous(lhs: Tuple[torch.Tensor, torch.Tensor],
                                              rhs: Tuple[torch.Tensor, torch.Tensor], out: torch.Tensor,
                                              m_indices: torch.Tensor, use_pdl: bool = False) -> None:
    """Grouped GEMM, contiguous layout (jit_kernels/m_grouped_gemm.py:55-127): lhs [m_sum, k] e4m3 with 1x128 scales
    [m_sum, k/128]; rhs [groups, n, k] e4m3 with 128x128 scales [groups, n/128, k/128]; out [m_sum, n] bf16;
    m_indices int32 [m_sum]: group of every row, constant over each 128-row block, -1 = skip the block."""
    lhs, lhs_scales = lhs
    rhs, rhs_scales = rhs
    m, k = lhs.shape
    num_groups, n, k_ = rhs.shape
    m_, n_ = out.shape
    assert m == m_ == m_indices.numel() and k == k_ and n == n_
    assert lhs_scales.shape == (m, (k + 127) // 128)
    assert rhs_scales.shape == (num_groups, (n + 127) // 128, (k + 127) // 128)
    assert lhs.dtype == torch.float8_e4m3fn and lhs_scales.dtype == torch.float32
    assert rhs.dtype == torch.float8_e4m3fn and rhs_scales.dtype == torch.float32
    assert out.dtype == torch.bfloat16 and m_indices.dtype == torch.int32
    assert lhs.is_contiguous() and rhs.is_contiguous() and out.is_contiguous() and m_indices.is_contiguous()
    assert m % get_m_alignment_for_contiguous_layout() == 0 and n % 128 == 0 and k % 128 == 0
    lhs_scales = get_col_major_tma_aligned_tensor(lhs_scales)
    assert rhs_scales.is_contiguous()
    if m == 0:
        return
    if not lhs.is_cuda:
        raise RuntimeError("m_grouped_gemm_fp8_fp8_bf16_nt_contiguous: tensors must be on a CUDA device (no CPU fallback)")
    with torch.cuda.device(lhs.device):
        _lib.check(
            _lib.lib().b200_fp8_gemm_grouped_nt(
                lhs.data_ptr(), lhs_scales.data_ptr(),
                lhs_scales.stride(1) if lhs_scales.shape[1] > 1 else get_tma_aligned_size(m, 4), rhs.data_ptr(),
                rhs_scales.data_ptr(), out.data_ptr(), m_indices.data_ptr(), 1, num_groups, m, n, k, int(bool(use_pdl)),
                _lib.current_stream()), "m_grouped_gemm_fp8_fp8_bf16_nt_contiguous")


def m_grouped_gemm_fp8_fp8_bf16_nt_masked(lhs: Tuple[torch.Tensor, torch.Tensor], rhs: Tuple[torch.Tensor, torch.Tensor],
                                          out: torch.Tensor, masked_m: torch.Tensor, expected_m: int,
                                          use_pdl: bool = False) -> None:
    """Grouped GEMM, masked layout (jit_kernels/m_grouped_gemm.py:130-215): lhs [groups, m_max, k] e4m3 with scales
    [groups, m_max, k/128]; rhs / scales as above; out [groups, m_max, n] bf16; masked_m int32 [groups]: rows of each
    group that are computed (rows past it are left untouched).  expected_m is the reference's tuning hint."""
    lhs, lhs_scales = lhs
    rhs, rhs_scales = rhs
    num_groups, m, k = lhs.shape
    num_groups_, n, k_ = rhs.shape
    num_groups__, m_, n_ = out.shape
    assert num_groups == num_groups_ == num_groups__ == masked_m.numel()
    assert m == m_ and n == n_ and k == k_
    assert expected_m > 0 and m > 0 and n > 0 and k > 0 and num_groups > 0
    assert lhs_scales.shape == (num_groups, m, (k + 127) // 128)
    assert rhs_scales.shape == (num_groups, (n + 127) // 128, (k + 127) // 128)
    assert lhs.dtype == torch.float8_e4m3fn and lhs_scales.dtype == torch.float32
    assert rhs.dtype == torch.float8_e4m3fn and rhs_scales.dtype == torch.float32
    assert out.dtype == torch.bfloat16 and masked_m.dtype == torch.int32
    assert lhs.is_contiguous() and rhs.is_contiguous() and out.is_contiguous() and masked_m.is_contiguous()
    assert n % 128 == 0 and k % 128 == 0
    assert num_groups == 1 or m % 128 == 0, "masked layout: m_max must be a multiple of the 128-row tile"
    lhs_scales = get_col_major_tma_aligned_tensor(lhs_scales)  # transposed per group
    assert rhs_scales.is_contiguous()
    if not lhs.is_cuda:
        raise RuntimeError("m_grouped_gemm_fp8_fp8_bf16_nt_masked: tensors must be on a CUDA device (no CPU fallback)")
    with torch.cuda.device(lhs.device):
        _lib.check(
            _lib.lib().b200_fp8_gemm_grouped_nt(
                lhs.data_ptr(), lhs_scales.data_ptr(),
                lhs_scales.stride(2) if lhs_scales.shape[2] > 1 else get_tma_aligned_size(m, 4), rhs.data_ptr(),
                rhs_scales.data_ptr(), out.data_ptr(), masked_m.data_ptr(), 2, num_groups, m, n, k, int(bool(use_pdl)),
                _lib.current_stream()), "m_grouped_gemm_fp8_fp8_bf16_nt_masked")


def m_grouped_gemm_fp8_fp8_bf16_nt_offset(lhs: Tuple[torch.Tensor, torch.Tensor], rhs: Tuple[torch.Tensor, torch.Tensor],
                                          out: torch.Tensor, m_offset: torch.Tensor, use_pdl: bool = False) -> None:
    """Grouped GEMM, offset layout (jit_kernels/m_grouped_gemm.py:225-310): lhs [m_sum, k] e4m3 with scales
    [m_sum, k/128]; rhs [groups, n, k] / scales [groups, n/128, k/128]; out [m_sum, n] bf16; m_offset int32
    [groups + 1]: group g owns rows [m_offset[g], m_offset[g + 1]) -- any row counts, no alignment padding."""
    lhs, lhs_scales = lhs
    rhs, rhs_scales = rhs
    m, k = lhs.shape
    num_groups, n, k_ = rhs.shape
    m_, n_ = out.shape
    assert m == m_ and m_offset.numel() - 1 == num_groups and k == k_ and n == n_
    assert lhs_scales.shape == (m, (k + 127) // 128)
    assert rhs_scales.shape == (num_groups, (n + 127) // 128, (k + 127) // 128)
    assert lhs.dtype == torch.float8_e4m3fn and lhs_scales.dtype == torch.float32
    assert rhs.dtype == torch.float8_e4m3fn and rhs_scales.dtype == torch.float32
    assert out.dtype == torch.bfloat16 and m_offset.dtype == torch.int32
    assert lhs.is_contiguous() and rhs.is_contiguous() and out.is_contiguous() and m_offset.is_contiguous()
    assert n % 128 == 0 and k % 128 == 0
    lhs_scales = get_col_major_tma_aligned_tensor(lhs_scales)
    assert rhs_scales.is_contiguous()
    if m == 0:
        return
    if not lhs.is_cuda:
        raise RuntimeError("m_grouped_gemm_fp8_fp8_bf16_nt_offset: tensors must be on a CUDA device (no CPU fallback)")
    with torch.cuda.device(lhs.device):
        _lib.check(
            _lib.lib().b200_fp8_gemm_grouped_nt(
                lhs.data_ptr(), lhs_scales.data_ptr(),
                lhs_scales.stride(1) if lhs_scales.shape[1] > 1 else get_tma_aligned_size(m, 4), rhs.data_ptr(),
                rhs_scales.data_ptr(), out.data_ptr(), m_offset.data_ptr(), 3, num_groups, m, n, k, int(bool(use_pdl)),
                _lib.current_stream()), "m_grouped_gemm_fp8_fp8_bf16_nt_offset")
