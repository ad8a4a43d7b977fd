"""Drop-in for the reference's ``flash_mla_fp8`` kernel module (3rdparty/flashmla-fp8/flash_mla_fp8/__init__.py:3-10,
flash_mla_interface.py:8-185): same function names, argument order, shapes and error behaviour (RuntimeError on bad
input), backed by the sm_100a C-ABI library instead of the sm_90a pybind extension ``flash_mla_fp8_cuda``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

PAGE_BLOCK_SIZE = 64
TILE_SCHEDULER_METADATA_SIZE = 8


def _require(cond: bool, msg: str) -> None:
    if not cond:
        raise RuntimeError(msg)  # the reference raises RuntimeError through TORCH_CHECK


def _cuda(t: torch.Tensor, name: str) -> None:
    _require(t.is_cuda, "%s must be on CUDA device" % name)  # flash_api.cpp:23 CHECK_DEVICE


REFERENCE_OVERHEAD_BLOCKS = 5  # get_mla_metadata.cu:16 fixed_overhead_num_blocks


def tuned_overhead_blocks(num_q_rows: int) -> int:
    """The scheduler's per-request fixed cost (in 64-token blocks) that balances the parts on B200: a request boundary
    costs the CTA-pair kernel (more than 32 q rows) about 24 blocks' worth of time and the SwapAB kernel about 12
    (measured; the reference's 5 is the sm_90 kernel's).  B200_MLA_SCHED_OVERHEAD overrides (sweeps)."""
    import os
    env = os.environ.get("B200_MLA_SCHED_OVERHEAD")
    if env:
        return int(env)
    return 24 if num_q_rows > 32 else 12


def tuned_num_sm_parts(num_q_rows: int) -> Optional[int]:
    """Scheduler parts for the kernel that will run, where that differs from the reference's rule (one part per SM and
    64 q rows, flash_api.cpp:41-43): 33-64 q rows run on CTA pairs, so half as many parts (one wave instead of two waves
    of short parts); None = the reference's rule."""
    import os
    env = os.environ.get("B200_MLA_PARTS")  # sweeps
    if env:
        return int(env)
    if 32 < num_q_rows <= 64:
        return max(1, int(_lib.lib().b200_sm_count()) // 2)
    return None


def get_mla_metadata(cache_seqlens: torch.Tensor, num_heads_per_head_k: int,
                     num_heads_k: int, fixed_overhead_blocks: Optional[int] = None,
                     num_sm_parts: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """flash_mla_interface.py:8-23 -> csrc/flash_api.cpp:27-63.
    Returns (tile_scheduler_metadata int32 [num_sm_parts, 8], num_splits int32 [batch + 1]).
    ``fixed_overhead_blocks`` (extension): the scheduler's per-request fixed cost; default = the reference's constant,
    i.e. the reference's schedule bit for bit; ``tuned_overhead_blocks(M)`` is what the B200 backend passes.
    ``num_sm_parts`` (extension): number of parts; default = the reference's rule, ``tuned_num_sm_parts(M)`` is the
    backend's."""
    _cuda(cache_seqlens, "seqlens_k")
    _require(cache_seqlens.dtype == torch.int32, "seqlens_k must have dtype int32")
    _require(cache_seqlens.is_contiguous(), "seqlens_k must be contiguous")
    if fixed_overhead_blocks is None:
        fixed_overhead_blocks = REFERENCE_OVERHEAD_BLOCKS
    L = _lib.lib()
    with torch.cuda.device(cache_seqlens.device):
        if num_sm_parts is None:
            num_sm_parts = L.b200_mla_num_sm_parts(num_heads_per_head_k, num_heads_k, L.b200_sm_count())
        _require(num_sm_parts > 0, "get_mla_metadata: no SM parts for num_heads_per_head_k=%d" % num_heads_per_head_k)
        batch = cache_seqlens.numel()
        meta = torch.empty(num_sm_parts, TILE_SCHEDULER_METADATA_SIZE, dtype=torch.int32, device=cache_seqlens.device)
        num_splits = torch.empty(batch + 1, dtype=torch.int32, device=cache_seqlens.device)
        _lib.check(L.b200_mla_get_metadata_ex(cache_seqlens.data_ptr(), batch, num_sm_parts, int(fixed_overhead_blocks),
                                              meta.data_ptr(), num_splits.data_ptr(), _lib.current_stream()),
                   "get_mla_metadata")
    return meta, num_splits


_merge_states: Dict[Tuple[int, int], torch.Tensor] = {}
_merge_retired: List[torch.Tensor] = []  # outgrown buffers: a captured graph may still point at them


def merge_state_ints(batch: int, s_q: int, h_q: int) -> int:
    """Entries of the int32 ``merge_state`` scratch of flash_mla_ckv_fp8_per_token."""
    return int(_lib.lib().b200_mla_merge_state_ints(batch, s_q, h_q))


def new_merge_state(batch: int, s_q: int, h_q: int, device) -> torch.Tensor:
    """A zeroed merge-state buffer for a caller that owns its own (one per concurrently running call)."""
    return torch.zeros(max(1, merge_state_ints(batch, s_q, h_q)), dtype=torch.int32, device=device)


def _default_merge_state(dev: torch.device, n_ints: int) -> Optional[torch.Tensor]:
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), torch.cuda.current_stream(dev).cuda_stream)
    buf = _merge_states.get(key)
    if buf is not None and buf.numel() >= n_ints:
        return buf
    if torch.cuda.is_current_stream_capturing():
        # a buffer zeroed inside the capture would only be zero after the first replay: leave this call on the
        # partials + combine path (correct, one more launch) rather than guess
        return None
    if buf is not None:
        _merge_retired.append(buf)
    buf = torch.zeros(max(n_ints, 1 << 14), dtype=torch.int32, device=dev)
    _merge_states[key] = buf
    return buf


def flash_mla_ckv_fp8_per_token(
    q_nope: torch.Tensor,
    q_rope: torch.Tensor,
    k_cache_lora: torch.Tensor,
    k_cache_rope: torch.Tensor,
    q_scale: torch.Tensor,
    k_scale: torch.Tensor,
    block_table: torch.Tensor,
    cache_seqlens: torch.Tensor,
    head_dim_v: int,
    tile_scheduler_metadata: torch.Tensor,
    num_splits: torch.Tensor,
    softmax_scale: Optional[float] = None,
    causal: bool = False,
    out: Optional[torch.Tensor] = None,
    lse: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    use_pdl: bool = True,
    merge_state=None,
) -> Tuple[torch.Tensor, torch.Tensor]:
    """flash_mla_interface.py:68-120 -> csrc/flash_api.cpp:216-397.

    ``merge_state``: int32 scratch (``merge_state_ints(b, s_q, h_q)`` entries, zero) with which the kernel for
    s_q * h_q > 32 merges requests cut over several parts itself instead of leaving them to a combine launch (see
    include/b200_decode.h).  Default: one zeroed buffer per (device, stream), kept by this module; a caller that runs
    the op on several streams at once or captures it into a graph before the first eager call passes its own;
    ``False`` keeps the reference's two launches (fp32 partials + combine).

    q_nope (b, s_q, h_q, 512) fp8/uint8; q_rope (b, s_q, h_q, 64) bf16; q_scale (b, s_q, h_q, 1) f32;
    k_cache_lora (num_blocks, 64, 1, 512) fp8/uint8; k_cache_rope (num_blocks, 64, 1, 64) bf16;
    k_scale (num_blocks, 64, 1, 1) f32; block_table (b, max_blocks) int32; cache_seqlens (b) int32.
    Returns (out (b, s_q, h_q, 512) bf16, softmax_lse (b, h_q, s_q) f32).
    ``out`` / ``lse`` / ``workspace`` may be passed in to keep the call allocation-free (CUDA graphs)."""
    for t, n in ((q_nope, "q_nope"), (q_rope, "q_rope"), (k_cache_lora, "k_cache_lora"), (k_cache_rope, "k_cache_rope"),
                 (q_scale, "q_scale"), (k_scale, "k_scale"), (block_table, "block_table"),
                 (cache_seqlens, "seqlens_k"), (tile_scheduler_metadata, "tile_scheduler_metadata"),
                 (num_splits, "num_splits")):
        _cuda(t, n)
    _require(q_nope.dim() == 4 and q_rope.dim() == 4, "q_nope / q_rope must be 4-D (b, s_q, h_q, d)")
    b, s_q, h_q, dv = q_nope.shape
    d_rope = q_rope.shape[-1]
    _require(head_dim_v == 512 and dv == 512 and d_rope == 64,
             "only head_dim_v=512 / rope=64 is supported")  # flash_api.cpp:252-255
    _require(q_nope.element_size() == 1, "q_nope must be fp8 (e4m3) / uint8")
    _require(k_cache_lora.element_size() == 1, "k_cache_lora must be fp8 (e4m3) / uint8")
    _require(q_rope.dtype == torch.bfloat16 and k_cache_rope.dtype == torch.bfloat16,
             "q_rope / k_cache_rope must be bfloat16")
    _require(q_scale.dtype == torch.float32 and k_scale.dtype == torch.float32, "scales must be float32")
    _require(block_table.dtype == torch.int32, "block_table must have dtype torch.int32")
    _require(cache_seqlens.dtype == torch.int32, "seqlens_k must have dtype int32")
    _require(block_table.stride(-1) == 1, "block_table must have contiguous last dimension")
    _require(k_cache_lora.dim() == 4 and k_cache_lora.shape[1] == PAGE_BLOCK_SIZE and k_cache_lora.shape[2] == 1,
             "k_cache_lora must be (num_blocks, 64, 1, 512)")
    _require(tuple(q_rope.shape[:3]) == (b, s_q, h_q) and q_scale.numel() == b * s_q * h_q, "q shape mismatch")
    num_blocks = k_cache_lora.shape[0]
    _require(k_cache_rope.shape[0] == num_blocks and k_scale.numel() == num_blocks * PAGE_BLOCK_SIZE,
             "k cache shape mismatch")
    for t, n in ((q_nope, "q_nope"), (q_rope, "q_rope"), (q_scale, "q_scale"), (k_cache_lora, "k_cache_lora"),
                 (k_cache_rope, "k_cache_rope"), (k_scale, "k_scale"), (cache_seqlens, "seqlens_k"),
                 (tile_scheduler_metadata, "tile_scheduler_metadata"), (num_splits, "num_splits")):
        _require(t.is_contiguous(), "%s must be contiguous" % n)
    _require(block_table.shape[0] == b and cache_seqlens.numel() == b, "batch size mismatch")
    _require(tile_scheduler_metadata.dtype == torch.int32 and tile_scheduler_metadata.dim() == 2
             and tile_scheduler_metadata.shape[1] == TILE_SCHEDULER_METADATA_SIZE,
             "tile_scheduler_metadata must be int32 (num_sm_parts, 8)")
    _require(num_splits.dtype == torch.int32 and num_splits.numel() == b + 1, "num_splits must be int32 (batch + 1)")
    if softmax_scale is None:
        softmax_scale = (dv + d_rope) ** (-0.5)  # flash_mla_interface.py:100-101
    num_sm_parts = tile_scheduler_metadata.shape[0]
    L = _lib.lib()
    dev = q_nope.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(b, s_q, h_q, dv, dtype=torch.bfloat16, device=dev)
        if lse is None:
            lse = torch.empty(b, h_q, s_q, dtype=torch.float32, device=dev)
        ws_bytes = L.b200_mla_workspace_bytes(b, s_q, h_q, num_sm_parts)
        if workspace is None:
            workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _require(workspace.numel() * workspace.element_size() >= ws_bytes, "workspace too small")
        n_ms = L.b200_mla_merge_state_ints(b, s_q, h_q)
        if merge_state is False:
            merge_state = None
        elif merge_state is None:
            merge_state = _default_merge_state(dev, n_ms)
        else:
            _cuda(merge_state, "merge_state")
            _require(merge_state.dtype == torch.int32 and merge_state.is_contiguous() and merge_state.numel() >= n_ms,
                     "merge_state must be contiguous int32 with at least %d entries" % n_ms)
        _lib.check(
            L.b200_mla_decode_ckv_fp8_per_token(
                q_nope.data_ptr(), q_rope.data_ptr(), q_scale.data_ptr(), k_cache_lora.data_ptr(),
                k_cache_rope.data_ptr(), k_scale.data_ptr(), num_blocks, block_table.data_ptr(), block_table.stride(0),
                cache_seqlens.data_ptr(), tile_scheduler_metadata.data_ptr(), num_splits.data_ptr(), num_sm_parts, b,
                s_q, h_q, float(softmax_scale), int(bool(causal)), out.data_ptr(), lse.data_ptr(),
                workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                merge_state.data_ptr() if merge_state is not None else None, int(bool(use_pdl)),
                _lib.current_stream()),
            "flash_mla_ckv_fp8_per_token")
    return out, lse


def quantize_ckv_per_token_head(tensor: torch.Tensor,
                                head_dim_v: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """flash_mla_interface.py:122-126 -> csrc/flash_api.cpp:399-460.
    tensor [..., 576] bf16 contiguous -> (nope uint8 [..., 512], scale f32 [..., 1], rope bf16 [..., 64])."""
    _cuda(tensor, "tensor")
    _require(tensor.dim() >= 2, "tensor must have at least 2 dims")
    _require(tensor.is_contiguous(), "tensor must be contiguous")
    _require(tensor.dtype == torch.bfloat16, "Unsupported tensor dtype")
    head_dim = tensor.shape[-1]
    base = tuple(tensor.shape[:-1])
    rows = tensor.numel() // head_dim
    dev = tensor.device
    nope = torch.empty(base + (head_dim_v,), dtype=torch.uint8, device=dev)
    scale = torch.empty(base + (1,), dtype=torch.float32, device=dev)
    rope = torch.empty(base + (head_dim - head_dim_v,), dtype=tensor.dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().b200_mla_quantize_q(tensor.data_ptr(), rows, head_dim, head_dim_v, nope.data_ptr(),
                                                  scale.data_ptr(), rope.data_ptr(), _lib.current_stream()),
                   "quantize_ckv_per_token_head")
    return nope, scale, rope


def quantize_and_cache_k(key: torch.Tensor, k_lora_cache: torch.Tensor, k_lora_scale_cache: torch.Tensor,
                         k_rope_cache: torch.Tensor, indices: torch.Tensor,
                         head_dim_v: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """flash_mla_interface.py:128-154 -> csrc/flash_api.cpp:462-532: quantise ``key`` rows and scatter them to the
    token slots ``indices`` of the three pool tensors, in place."""
    _cuda(key, "key")
    for t, n in ((k_lora_cache, "k_lora_cache"), (k_lora_scale_cache, "k_lora_scale_cache"),
                 (k_rope_cache, "k_rope_cache"), (indices, "indices")):
        _cuda(t, n)
        _require(t.is_contiguous(), "%s must be contiguous" % n)
    _require(indices.dtype == torch.int32, "indices must have dtype int32")
    _require(key.dtype == torch.bfloat16, "Unsupported key dtype")
    head_dim = key.shape[-1]
    n = key.numel() // head_dim
    # rows may be a strided view (e.g. the kv columns of the fused qkv_a projection output): [n, (1,) head_dim] with a
    # contiguous last dimension and one row stride
    _require(key.stride(-1) == 1 and key.dim() in (2, 3) and (key.dim() == 2 or key.shape[1] == 1),
             "key must be [n, head_dim] or [n, 1, head_dim] with a contiguous last dimension")
    key_row_stride = key.stride(0) if n > 1 else head_dim
    _require(indices.numel() == n, "indices size must match key batch size")
    _require(k_lora_cache.element_size() == 1 and k_lora_cache.shape[-1] == head_dim_v, "k_lora_cache shape/dtype")
    _require(k_lora_scale_cache.dtype == torch.float32, "k_lora_scale_cache must be float32")
    _require(k_rope_cache.dtype == key.dtype and k_rope_cache.shape[-1] == head_dim - head_dim_v, "k_rope_cache shape")
    with torch.cuda.device(key.device):
        _lib.check(
            _lib.lib().b200_mla_quantize_and_cache_k(key.data_ptr(), key_row_stride, n, head_dim, head_dim_v,
                                                     indices.data_ptr(), k_lora_cache.data_ptr(),
                                                     k_lora_scale_cache.data_ptr(), k_rope_cache.data_ptr(),
                                                     _lib.current_stream()), "quantize_and_cache_k")
    return k_lora_cache, k_lora_scale_cache, k_rope_cache


def dequantize_ckv_fused_indexed(k_lora_cache: torch.Tensor, k_rope_cache: torch.Tensor, k_scale_cache: torch.Tensor,
                                 indices: torch.Tensor, max_num_tokens: int = -1) -> Tuple[torch.Tensor, torch.Tensor]:
    """flash_mla_interface.py:156-185 -> csrc/flash_api.cpp:534-632: gather ``indices`` and de-quantise to bf16.
    Returns (k_lora_deq [n, 1, 512], k_rope_deq [n, 1, 64])."""
    for t, n in ((k_lora_cache, "k_lora_cache"), (k_rope_cache, "k_rope_cache"), (k_scale_cache, "k_scale_cache"),
                 (indices, "indices")):
        _cuda(t, n)
        _require(t.is_contiguous(), "%s must be contiguous" % n)
    _require(indices.dtype == torch.int32, "indices must be int32")
    _require(k_lora_cache.element_size() == 1, "k_lora_cache must be fp8 / uint8")
    _require(k_rope_cache.dtype == torch.bfloat16, "k_rope_cache must be bfloat16")
    _require(k_scale_cache.dtype == torch.float32, "k_scale_cache must be float32")
    n = indices.shape[0]
    dev = k_lora_cache.device
    lora = torch.empty(n, 1, k_lora_cache.shape[-1], dtype=k_rope_cache.dtype, device=dev)
    rope = torch.empty(n, 1, k_rope_cache.shape[-1], dtype=k_rope_cache.dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(
            _lib.lib().b200_mla_dequantize_indexed(k_lora_cache.data_ptr(), k_rope_cache.data_ptr(),
                                                   k_scale_cache.data_ptr(), indices.data_ptr(), n,
                                                   int(max_num_tokens), lora.data_ptr(), rope.data_ptr(),
                                                   _lib.current_stream()), "dequantize_ckv_fused_indexed")
    return lora, rope


def flash_mla_with_kvcache(q: torch.Tensor, k_cache: torch.Tensor, block_table: torch.Tensor,
                           cache_seqlens: torch.Tensor, head_dim_v: int, tile_scheduler_metadata: torch.Tensor,
                           num_splits: torch.Tensor, softmax_scale: Optional[float] = None, causal: bool = False,
                           out: Optional[torch.Tensor] = None, lse: Optional[torch.Tensor] = None,
                           workspace: Optional[torch.Tensor] = None,
                           use_pdl: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """bf16-KV entry point of the reference module (flash_mla_interface.py:26-66 -> csrc/flash_api.cpp:65-214), also the
    API of ``flash_mla_swap`` (flashmla-swap/flash_mla_swap/flash_mla_interface.py:26-66).

    q (b, s_q, h_q, 576) bf16; k_cache (num_blocks, 64, 1, 576) bf16; block_table (b, max_blocks) int32;
    cache_seqlens (b) int32.  Returns (out (b, s_q, h_q, head_dim_v) bf16, softmax_lse (b, h_q, s_q) f32).
    Runs the bf16 kernel (csrc/mla_decode_bf16.cu): nothing is quantised on this entry."""
    for t, n in ((q, "q"), (k_cache, "k_cache"), (block_table, "block_table"), (cache_seqlens, "seqlens_k"),
                 (tile_scheduler_metadata, "tile_scheduler_metadata"), (num_splits, "num_splits")):
        _cuda(t, n)
    _require(q.dtype == torch.bfloat16, "q must be bfloat16")  # flash_api.cpp:84
    _require(k_cache.dtype == q.dtype, "query and key must have the same dtype")
    _require(q.dim() == 4, "q must be 4-D (b, s_q, h_q, d)")
    b, s_q, h_q, d = q.shape
    _require(head_dim_v == 512 and d == 576, "only head_dim=576 / head_dim_v=512 is supported")  # flash_api.cpp:103-105
    _require(k_cache.dim() == 4 and k_cache.shape[1] == PAGE_BLOCK_SIZE and k_cache.shape[2] == 1
             and k_cache.shape[3] == d, "k_cache must be (num_blocks, 64, 1, 576)")
    _require(q.stride(-1) == 1 and k_cache.stride(-1) == 1, "q / k_cache must have contiguous last dimension")
    _require(block_table.dtype == torch.int32, "block_table must have dtype torch.int32")
    _require(block_table.stride(-1) == 1, "block_table must have contiguous last dimension")
    _require(cache_seqlens.dtype == torch.int32, "seqlens_k must have dtype int32")
    _require(block_table.shape[0] == b and cache_seqlens.numel() == b, "batch size mismatch")
    for t, n in ((cache_seqlens, "seqlens_k"), (tile_scheduler_metadata, "tile_scheduler_metadata"),
                 (num_splits, "num_splits")):
        _require(t.is_contiguous(), "%s must be contiguous" % n)
    _require(tile_scheduler_metadata.dtype == torch.int32 and tile_scheduler_metadata.dim() == 2
             and tile_scheduler_metadata.shape[1] == TILE_SCHEDULER_METADATA_SIZE,
             "tile_scheduler_metadata must be int32 (num_sm_parts, 8)")
    _require(num_splits.dtype == torch.int32 and num_splits.numel() == b + 1, "num_splits must be int32 (batch + 1)")
    if softmax_scale is None:
        softmax_scale = d ** (-0.5)  # flash_mla_interface.py:52-53
    q = q.contiguous()
    k_cache = k_cache.contiguous()
    num_blocks = k_cache.shape[0]
    num_sm_parts = tile_scheduler_metadata.shape[0]
    L = _lib.lib()
    dev = q.device
    with torch.cuda.device(dev):
        if out is None:
            out = torch.empty(b, s_q, h_q, head_dim_v, dtype=torch.bfloat16, device=dev)
        if lse is None:
            lse = torch.empty(b, h_q, s_q, dtype=torch.float32, device=dev)
        ws_bytes = L.b200_mla_workspace_bytes(b, s_q, h_q, num_sm_parts)
        if workspace is None:
            workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _require(workspace.numel() * workspace.element_size() >= ws_bytes, "workspace too small")
        _lib.check(
            L.b200_mla_decode_bf16(q.data_ptr(), k_cache.data_ptr(), num_blocks, block_table.data_ptr(),
                                   block_table.stride(0), cache_seqlens.data_ptr(), tile_scheduler_metadata.data_ptr(),
                                   num_splits.data_ptr(), num_sm_parts, b, s_q, h_q, float(softmax_scale),
                                   int(bool(causal)), out.data_ptr(), lse.data_ptr(), workspace.data_ptr(),
                                   workspace.numel() * workspace.element_size(), int(bool(use_pdl)),
                                   _lib.current_stream()), "flash_mla_with_kvcache")
    return out, lse
