"""Fused decode prologue for the per-token-FP8 MLA pool.

The reference's absorbed-MLA decode runs, per layer, ``rotary_emb(positions, q_pe, k_pe, fused_set_kv_buffer_arg=...,
output_q_rope=...)`` (python/sglang/srt/models/deepseek_v2.py:843-857 -> layers/rotary_embedding.py:198-213 ->
flashinfer ``apply_rope_with_cos_sin_cache_inplace``), then ``set_kv_buffer`` -> ``quantize_and_cache_k``
(mem_cache/memory_pool.py:854-871) and ``quantize_ckv_per_token_head`` on a ``.contiguous()`` copy of Q
(layers/attention/flashmla_backend.py:205-206).  Its ``fused_set_kv_buffer_arg`` hook only covers the bf16 pool
(models/utils.py:52-82 slices one bf16 tensor); for the FP8 pool all three stay separate launches.  Here they are one
kernel (``b200_mla_rope_quantize_fused``) whose outputs are bit-identical to that sequence; the names below follow the
reference's hook so a model file changes one call.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib


@dataclass
class FusedSetKVBufferArg:
    """flashinfer/rope.py:28-49 extended to the three tensors of the per_token_head pool.

    value: the new tokens' normed latent, bf16 ``[tokens, 512]`` (a view of ``latent_cache`` is fine);
    k_buffer: rope cache bf16 ``[slots, 64]``; v_buffer: latent cache uint8/e4m3 ``[slots, 512]``;
    scale_buffer: f32 ``[slots]`` (any shape with that many elements); cache_loc: ``[tokens]`` int32/int64 slots.
    ``k_scale`` / ``v_scale`` (per-tensor scales) are not supported, as in the reference (rope.py:1217-1218)."""

    value: torch.Tensor
    k_buffer: torch.Tensor
    v_buffer: torch.Tensor
    scale_buffer: torch.Tensor
    k_scale: Optional[float] = None
    v_scale: Optional[float] = None
    cache_loc: Optional[torch.Tensor] = None


def create_fused_set_kv_buffer_arg(value: torch.Tensor, layer, forward_batch) -> FusedSetKVBufferArg:
    """models/utils.py:52-82 for ``MLATokenToKVPool(quant_method="per_token_head")``."""
    pool = forward_batch.token_to_kv_pool
    if getattr(pool, "quant_method", None) != "per_token_head":
        raise ValueError("create_fused_set_kv_buffer_arg: the fused prologue writes the per_token_head pool layout")
    k_lora, k_scale, k_rope = pool.get_key_buffer(layer.layer_id)
    return FusedSetKVBufferArg(value=value, k_buffer=k_rope.view(k_rope.shape[0], -1),
                               v_buffer=k_lora.view(k_lora.shape[0], -1), scale_buffer=k_scale.view(-1),
                               cache_loc=forward_batch.out_cache_loc)


def _last_dim_contiguous(t: torch.Tensor, name: str) -> None:
    if t.stride(-1) != 1:
        raise RuntimeError("%s: last dimension must be contiguous" % name)


def fused_rope_quantize_set_kv(positions: torch.Tensor, q_nope: torch.Tensor, q_pe: torch.Tensor, k_pe: torch.Tensor,
                               cos_sin_cache: torch.Tensor, is_neox: bool,
                               fused_set_kv_buffer_arg: Optional[FusedSetKVBufferArg] = None
                               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Rotate, quantise, and (with ``fused_set_kv_buffer_arg``) cache, in one launch.

    positions ``[tokens]``; q_nope bf16 ``[tokens, heads, 512]`` (the ``w_kc`` bmm output, any token/head strides);
    q_pe bf16 ``[tokens, heads, 64]`` before rotation (e.g. the split view of q); k_pe bf16 ``[tokens, 64]`` or
    ``[tokens, 1, 64]`` before rotation -- with ``fused_set_kv_buffer_arg`` it must be the tail of the same 576-wide rows
    whose head is ``fused_set_kv_buffer_arg.value`` (``latent_cache``, deepseek_v2.py:826-839); cos_sin_cache f32
    ``[max_position, 64]``.  Returns what ``quantize_ckv_per_token_head`` returns for the rotated Q:
    (q_nope uint8 ``[tokens, heads, 512]``, q_scale f32 ``[tokens, heads, 1]``, q_rope bf16 ``[tokens, heads, 64]``)."""
    for t, n in ((positions, "positions"), (q_nope, "q_nope"), (q_pe, "q_pe"), (cos_sin_cache, "cos_sin_cache")):
        if not t.is_cuda:
            raise RuntimeError("fused_rope_quantize_set_kv: %s must be a CUDA tensor (no CPU fallback)" % n)
    if cos_sin_cache.dtype != torch.float32:
        raise ValueError("cos_sin_cache should be float32")  # rope.py:1212-1213
    if cos_sin_cache.dim() != 2 or cos_sin_cache.shape[1] != 64 or not cos_sin_cache.is_contiguous():
        raise RuntimeError("fused_rope_quantize_set_kv: cos_sin_cache must be contiguous [max_position, 64]")
    if q_nope.dtype != torch.bfloat16 or q_pe.dtype != torch.bfloat16:
        raise RuntimeError("fused_rope_quantize_set_kv: q_nope / q_pe must be bfloat16")
    if q_nope.dim() != 3 or q_pe.dim() != 3 or q_nope.shape[-1] != 512 or q_pe.shape[-1] != 64 \
            or q_nope.shape[:2] != q_pe.shape[:2]:
        raise RuntimeError("fused_rope_quantize_set_kv: q_nope [tokens, heads, 512] / q_pe [tokens, heads, 64] expected")
    _last_dim_contiguous(q_nope, "q_nope")
    _last_dim_contiguous(q_pe, "q_pe")
    tokens, heads = q_nope.shape[:2]
    if positions.numel() != tokens:
        raise RuntimeError("fused_rope_quantize_set_kv: positions must have one entry per token")
    pos_ids = positions.reshape(-1).to(torch.int64).contiguous()  # rope.py:1237
    dev = q_nope.device
    out_nope = torch.empty((tokens, heads, 512), dtype=torch.uint8, device=dev)
    out_scale = torch.empty((tokens, heads, 1), dtype=torch.float32, device=dev)
    out_rope = torch.empty((tokens, heads, 64), dtype=torch.bfloat16, device=dev)
    kv_ptr, kv_stride, loc_ptr, lora_ptr, scale_ptr, rope_ptr = None, 0, None, None, None, None
    a = fused_set_kv_buffer_arg
    if a is not None:
        if a.k_scale is not None or a.v_scale is not None:
            raise ValueError("k_scale/v_scale are not supported yet")  # rope.py:1217-1218
        if a.cache_loc is None:
            raise ValueError("fused_set_kv_buffer_arg.cache_loc is required")  # rope.py:1219-1220
        if a.cache_loc.dtype not in (torch.int32, torch.int64):
            raise ValueError(f"cache_loc must be int32 or int64, got {a.cache_loc.dtype}")  # rope.py:1221-1222
        value = a.value.reshape(tokens, -1) if a.value.dim() != 2 else a.value
        k_pe2 = k_pe.view(tokens, 64) if k_pe.dim() == 3 else k_pe
        _last_dim_contiguous(value, "fused_set_kv_buffer_arg.value")
        _last_dim_contiguous(k_pe2, "k_pe")
        if value.shape != (tokens, 512) or k_pe2.shape != (tokens, 64) or value.dtype != torch.bfloat16 \
                or k_pe2.dtype != torch.bfloat16:
            raise RuntimeError("fused_rope_quantize_set_kv: value [tokens, 512] / k_pe [tokens, 64] bfloat16 expected")
        # one 576-wide row per token: latent | rope (the layout of latent_cache, deepseek_v2.py:826-839)
        if k_pe2.data_ptr() != value.data_ptr() + 512 * 2 or (tokens > 1 and k_pe2.stride(0) != value.stride(0)):
            raise RuntimeError("fused_rope_quantize_set_kv: value and k_pe must be the [:512] / [512:] views of the "
                               "same [tokens, 576] rows")
        if a.v_buffer.element_size() != 1 or a.v_buffer.shape[-1] != 512 or a.k_buffer.dtype != torch.bfloat16 \
                or a.k_buffer.shape[-1] != 64 or a.scale_buffer.dtype != torch.float32 \
                or not (a.v_buffer.is_contiguous() and a.k_buffer.is_contiguous() and a.scale_buffer.is_contiguous()):
            raise RuntimeError("fused_rope_quantize_set_kv: pool buffers must be contiguous uint8 [slots, 512] / "
                               "bfloat16 [slots, 64] / float32 [slots]")
        if a.cache_loc.numel() != tokens:
            raise RuntimeError("fused_rope_quantize_set_kv: cache_loc must have one slot per token")
        cache_loc = a.cache_loc.reshape(-1).to(torch.int32).contiguous()
        kv_ptr, kv_stride = value.data_ptr(), (value.stride(0) if tokens > 1 else 576)
        loc_ptr, lora_ptr, scale_ptr, rope_ptr = (cache_loc.data_ptr(), a.v_buffer.data_ptr(),
                                                  a.scale_buffer.data_ptr(), a.k_buffer.data_ptr())
    if tokens == 0:
        return out_nope, out_scale, out_rope
    # a size-1 dimension may report any stride: hand the kernel the canonical one
    qn_sh = q_nope.stride(1) if heads > 1 else 512
    qn_st = q_nope.stride(0) if tokens > 1 else heads * qn_sh
    qp_sh = q_pe.stride(1) if heads > 1 else 64
    qp_st = q_pe.stride(0) if tokens > 1 else heads * qp_sh
    with torch.cuda.device(dev):
        _lib.check(
            _lib.lib().b200_mla_rope_quantize_fused(
                q_nope.data_ptr(), qn_st, qn_sh, q_pe.data_ptr(), qp_st, qp_sh,
                kv_ptr, kv_stride, pos_ids.data_ptr(), cos_sin_cache.data_ptr(), 64, int(bool(is_neox)), tokens, heads,
                out_nope.data_ptr(), out_scale.data_ptr(), out_rope.data_ptr(), loc_ptr, lora_ptr, scale_ptr, rope_ptr,
                _lib.current_stream()), "fused_rope_quantize_set_kv")
    return out_nope, out_scale, out_rope
