"""Oracle: the KV-pool side of the path -- MLATokenToKVPool per_token_head layout, its torch fall-back quantiser,
the page allocator's block table and the torch-native SDPA decode (TEST INFRASTRUCTURE ONLY).

numpy restatements of reference Python (paths relative to /root/reference/python/sglang/srt); each is pinned by a
golden fixture generated from the reference's own code (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

from .fp8 import e4m3_to_f32, f32_to_e4m3, round_to_bf16

PAGE_SIZE = 64


def pool_fallback_quantize(key_f32: np.ndarray, kv_lora_rank: int = 512):
    """mem_cache/memory_pool.py:872-880 (non-sm90 branch of MLATokenToKVPool.set_kv_buffer):
    scale = amax(|k_lora|).clamp(1e-26) / 448; k_lora = (k_lora / scale) -> e4m3; k_rope = (k_rope / scale) -> bf16.
    Differs from the CUDA quantiser (cache_kernels.cu:76-79: + 1e-12, multiply by 1/scale) by at most one rounding."""
    key = np.asarray(key_f32, dtype=np.float32)
    lora, rope = key[..., :kv_lora_rank], key[..., kv_lora_rank:]
    amax = np.maximum(np.max(np.abs(lora), axis=-1, keepdims=True), np.float32(1e-26)).astype(np.float32)
    scale = (amax / np.float32(448.0)).astype(np.float32)
    codes = f32_to_e4m3((lora / scale).astype(np.float32))
    rope_q = round_to_bf16((rope / scale).astype(np.float32))
    return codes, scale, rope_q


def pool_fallback_dequantize(k_lora_codes, k_scale, k_rope_f32, indices):
    """mem_cache/memory_pool.py:826-831 (non-sm90 branch of get_key_split_contiguous)."""
    idx = np.asarray(indices, dtype=np.int64)
    lora = e4m3_to_f32(np.asarray(k_lora_codes)[idx])
    scale = np.asarray(k_scale, dtype=np.float32)[idx]
    rope = np.asarray(k_rope_f32, dtype=np.float32)[idx]
    return round_to_bf16((lora * scale).astype(np.float32)), round_to_bf16((rope * scale).astype(np.float32))


class KVAllocator:
    """mem_cache/allocator.py:19-102, 169-173: req_to_page int32 [max_bs, ceil(ctx/page)] block table; pages are
    handed out in increasing order starting at 1 (page 0 is the padding page)."""

    def __init__(self, size: int, max_batch_size: int, max_context_len: int, page_size: int = PAGE_SIZE):
        self.page_size = page_size
        self.max_page_num = (max_context_len + page_size - 1) // page_size
        self.req_to_page = np.zeros((max_batch_size, self.max_page_num), dtype=np.int32)
        self.free_slots = list(range(1, size // page_size))

    def alloc(self, req_pool_index: int, need_size: int, alloced_len: int):
        ps = self.page_size
        page_offset = alloced_len % ps
        page_num = (alloced_len + ps - 1) // ps
        last_page_remain = page_num * ps - alloced_len
        last_page_id = int(self.req_to_page[req_pool_index, page_num - 1])  # page_num 0 -> index -1, as upstream
        kv_loc = last_page_id * ps + page_offset + np.arange(0, min(last_page_remain, need_size), dtype=np.int32)
        if last_page_remain >= need_size:
            return kv_loc.astype(np.int32)
        remain = need_size - last_page_remain
        need_pages = (remain + ps - 1) // ps
        if need_pages > len(self.free_slots) or page_num + need_pages > self.max_page_num:
            return None
        new_pages = np.array(self.free_slots[:need_pages], dtype=np.int32)
        self.free_slots = self.free_slots[need_pages:]
        self.req_to_page[req_pool_index, page_num:page_num + need_pages] = new_pages
        kv_loc1 = (new_pages[:, None] * ps + np.arange(ps, dtype=np.int32)[None, :]).reshape(-1)[:remain]
        return np.concatenate([kv_loc, kv_loc1]).astype(np.int32)


def block_table_gather(req_to_page: np.ndarray, req_pool_indices) -> np.ndarray:
    """layers/attention/flashmla_backend.py:259,333,377: block_table = req_to_page[req_pool_indices] (bit-exact)."""
    return np.asarray(req_to_page)[np.asarray(req_pool_indices, dtype=np.int64)].astype(np.int32)


def native_sdpa_decode(q, k_cache, req_to_token, seq_lens, head_dim_v: int = 512, scaling=None):
    """layers/attention/torch_native_backend.py:275-343 with the absorbed-MLA use of it (:472-528): per sequence,
    gather k = k_cache[req_to_token[i, :len]], v = k[..., :512], softmax(q k^T * scale) v, one KV head shared by all
    q heads (enable_gqa).  float64 math, output rounded to bf16 like the torch bf16 SDPA result dtype."""
    q = np.asarray(q, dtype=np.float64)
    kc = np.asarray(k_cache, dtype=np.float64)
    b, h, d = q.shape
    if scaling is None:
        scaling = d ** -0.5
    out = np.zeros((b, h, head_dim_v))
    for i in range(b):
        n = int(seq_lens[i])
        k = kc[np.asarray(req_to_token[i, :n], dtype=np.int64), 0, :]
        logits = (q[i] @ k.T) * scaling
        logits -= logits.max(axis=1, keepdims=True)
        p = np.exp(logits)
        p /= p.sum(axis=1, keepdims=True)
        out[i] = p @ k[:, :head_dim_v]
    return out.astype(np.float32)
