"""Attention plug-in for the quantised decode path: ``FlashMLABackend`` with the reference's method names, argument
order and metadata handling (python/sglang/srt/layers/attention/flashmla_backend.py:49-408), an
``AttentionBackend`` base with the reference's ``forward`` dispatch (base_attn_backend.py:15-127), and the two
pool-side classes the backend reads -- ``MLATokenToKVPool`` in the per-token-FP8 layout
(mem_cache/memory_pool.py:591-882) and ``KVAllocator`` (mem_cache/allocator.py:19-102).

In the reference the backend is imported only when ``is_sm90_supported()`` (model_executor/attn_initializer.py:25-29)
and inherits its prefill from FlashInferMLAAttnBackend.  Here the class duck-types ``model_runner`` / ``forward_batch``
/ ``layer`` (so it can be registered in the reference's AttnInitializer -- see INTEGRATION.md -- or driven directly by
tests and bench.py), has the reference's constructor, and covers DECODE / IDLE / TARGET_VERIFY / DRAFT_EXTEND itself.
EXTEND / MIXED (prefill) is outside the hot path: it is delegated to a *prefill backend* -- inside the reference that is
an instance of its FlashInferMLAAttnBackend, built with the same arguments (``set_prefill_backend_factory``; done by
``install.patch_attention_initializer`` when that class imports) -- and raises when none is registered.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from enum import IntEnum, auto
from typing import Optional

import torch

from . import flash_mla_fp8

PAGE_SIZE = 64  # flashmla_backend.py:25

# ---- the two things the reference's class takes from its surroundings rather than from its arguments ----------------
_attention_tp_size_getter = None   # reference: sglang.srt.layers.dp_attention.get_attention_tp_size (flashmla_backend.py:11-13,67-69)
_prefill_backend_factory = None    # reference: the FlashInferMLAAttnBackend base class (flashmla_backend.py:49,59)


def set_attention_tp_size_getter(fn) -> None:
    """Inject ``get_attention_tp_size`` (tests, bench.py, or an embedding runtime); None restores the default lookup."""
    global _attention_tp_size_getter
    _attention_tp_size_getter = fn


def get_attention_tp_size() -> int:
    """The reference resolves heads per rank with dp_attention.get_attention_tp_size() (flashmla_backend.py:67-69), so
    ``FlashMLABackend(model_runner)`` -- the call attn_initializer.py:395 makes -- must work under TP > 1 without an
    extra argument: injected getter first, then the reference's own function when sglang is importable, then 1."""
    if _attention_tp_size_getter is not None:
        return int(_attention_tp_size_getter())
    import sys
    dp = sys.modules.get("sglang.srt.layers.dp_attention")
    if dp is None:
        try:
            import importlib
            dp = importlib.import_module("sglang.srt.layers.dp_attention")
        except Exception:
            return 1
    return int(dp.get_attention_tp_size())


def set_prefill_backend_factory(fn) -> None:
    """``fn(model_runner, skip_prefill, kv_indptr_buf, q_indptr_decode_buf)`` -> an object with the AttentionBackend
    methods that serves EXTEND / MIXED batches (the reference's FlashInferMLAAttnBackend); None unregisters."""
    global _prefill_backend_factory
    _prefill_backend_factory = fn


class ForwardMode(IntEnum):
    """model_executor/forward_batch_info.py:61-101 (same members and predicates)."""
    EXTEND = auto()
    DECODE = auto()
    MIXED = auto()
    IDLE = auto()
    TARGET_VERIFY = auto()
    DRAFT_EXTEND = auto()
    DUMMY_FIRST = auto()

    def is_extend(self):
        return self in (ForwardMode.EXTEND, ForwardMode.MIXED, ForwardMode.DRAFT_EXTEND, ForwardMode.TARGET_VERIFY)

    def is_decode(self):
        return self == ForwardMode.DECODE

    def is_idle(self):
        return self == ForwardMode.IDLE

    def is_decode_or_idle(self):
        return self == ForwardMode.DECODE or self == ForwardMode.IDLE

    def is_target_verify(self):
        return self == ForwardMode.TARGET_VERIFY

    def is_draft_extend(self):
        return self == ForwardMode.DRAFT_EXTEND

    def is_cuda_graph(self):
        return self in (ForwardMode.DECODE, ForwardMode.TARGET_VERIFY, ForwardMode.IDLE)


class AttentionBackend(ABC):
    """base_attn_backend.py:15-127."""

    @abstractmethod
    def init_forward_metadata(self, forward_batch):
        raise NotImplementedError()

    def init_cuda_graph_state(self, max_bs: int):
        raise NotImplementedError()

    def init_forward_metadata_capture_cuda_graph(self, bs, num_tokens, req_pool_indices, seq_lens, forward_mode,
                                                 spec_info):
        raise NotImplementedError()

    def init_forward_metadata_replay_cuda_graph(self, bs, req_pool_indices, seq_lens, seq_lens_sum, forward_mode,
                                                spec_info):
        raise NotImplementedError()

    def get_cuda_graph_seq_len_fill_value(self):
        raise NotImplementedError()

    def forward(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, **kwargs):
        if forward_batch.forward_mode.is_decode():
            return self.forward_decode(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache, **kwargs)
        return self.forward_extend(q, k, v, layer, forward_batch, save_kv_cache=save_kv_cache, **kwargs)

    def forward_decode(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, **kwargs):
        raise NotImplementedError()

    def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache: bool = True, **kwargs):
        raise NotImplementedError()


class KVAllocator:
    """mem_cache/allocator.py:19-102,169-173: owns ``req_to_page`` (int32 [max_bs, ceil(ctx/page)]) and hands out
    pages in increasing order from 1 (page 0 is the padding page)."""

    def __init__(self, size: int, device, max_batch_size: int, max_context_len: int, page_size: int = PAGE_SIZE):
        self.size, self.device, self.page_size = size, device, page_size
        self.max_page_num = (max_context_len + page_size - 1) // page_size
        self.req_to_page = torch.zeros((max_batch_size, self.max_page_num), dtype=torch.int32, device=device)
        self.clear()

    def clear(self):
        self.free_slots = torch.arange(1, self.size // self.page_size, dtype=torch.int32)

    def available_size(self):
        return len(self.free_slots)

    def alloc(self, req_pool_index: int, need_size: int, alloced_len: int):
        ps = self.page_size
        page_offset = alloced_len % ps
        page_num = (alloced_len + ps - 1) // ps
        last_page_remain = page_num * ps - alloced_len
        last_page_id = int(self.req_to_page[req_pool_index, page_num - 1].item())
        kv_loc = last_page_id * ps + page_offset + torch.arange(0, min(last_page_remain, need_size), dtype=torch.int32)
        if last_page_remain >= need_size:
            return kv_loc.to(self.device, non_blocking=True)
        remain = need_size - last_page_remain
        need_pages = (remain + ps - 1) // ps
        if need_pages > len(self.free_slots) or page_num + need_pages > self.max_page_num:
            return None
        new_pages = self.free_slots[:need_pages]
        self.free_slots = self.free_slots[need_pages:]
        self.req_to_page[req_pool_index, page_num:page_num + need_pages] = new_pages.to(self.device)
        kv_loc1 = (new_pages.unsqueeze(1) * ps + torch.arange(0, ps, dtype=torch.int32)).flatten()[:remain]
        return torch.concat([kv_loc, kv_loc1]).to(self.device, non_blocking=True)


class MLATokenToKVPool:
    """Per-layer MLA KV pool (mem_cache/memory_pool.py:604-880), both layouts the FlashMLA backend reads:

    * ``quant_method="per_token_head"`` (memory_pool.py:635-649): for every layer a tuple (k_lora uint8/e4m3
      [size+page, 1, 512], k_scale f32 [size+page, 1, 1], k_rope bf16 [size+page, 1, 64]) -- 644 bytes per token per
      layer;
    * ``quant_method="none"`` (memory_pool.py:650-660): one bf16 tensor [size+page, 1, 576] per layer -- 1152 bytes
      per token per layer.

    Page 0 absorbs padded writes."""

    def __init__(self, size: int, model_dtype: torch.dtype, kv_lora_rank: int, qk_rope_head_dim: int, layer_num: int,
                 device, page_size: int = PAGE_SIZE, quant_method: str = "per_token_head"):
        assert model_dtype == torch.bfloat16 and kv_lora_rank == 512 and qk_rope_head_dim == 64
        if quant_method not in ("per_token_head", "none"):
            raise ValueError("unknown kv cache quant method %r" % (quant_method,))
        self.size, self.page_size, self.device = size, page_size, device
        self.kv_lora_rank, self.qk_rope_head_dim, self.model_dtype = kv_lora_rank, qk_rope_head_dim, model_dtype
        self.quant_method = quant_method
        n = size + page_size
        if quant_method == "per_token_head":
            self.dtype = torch.float8_e4m3fn
            self.store_dtype = torch.uint8
            self.kv_buffer = [(torch.zeros((n, 1, kv_lora_rank), dtype=torch.uint8, device=device),
                               torch.zeros((n, 1, 1), dtype=torch.float32, device=device),
                               torch.zeros((n, 1, qk_rope_head_dim), dtype=model_dtype, device=device))
                              for _ in range(layer_num)]
        else:
            self.dtype = self.store_dtype = model_dtype
            self.kv_buffer = [torch.zeros((n, 1, kv_lora_rank + qk_rope_head_dim), dtype=model_dtype, device=device)
                              for _ in range(layer_num)]

    def get_key_buffer(self, layer_id: int):
        return self.kv_buffer[layer_id]  # memory_pool.py:810-811

    def get_value_buffer(self, layer_id: int):
        if self.quant_method == "per_token_head":
            return self.kv_buffer[layer_id][:2]  # memory_pool.py:844-845
        return self.kv_buffer[layer_id][..., : self.kv_lora_rank]  # memory_pool.py:846-849

    def set_kv_buffer(self, layer, loc: torch.Tensor, cache_k: torch.Tensor, cache_v=None):
        """memory_pool.py:854-871: quantise + scatter in one kernel (per-token FP8) or a plain indexed store."""
        if self.quant_method == "per_token_head":
            k_lora, k_scale, k_rope = self.kv_buffer[layer.layer_id]
            flash_mla_fp8.quantize_and_cache_k(key=cache_k.contiguous(), k_lora_cache=k_lora,
                                               k_lora_scale_cache=k_scale, k_rope_cache=k_rope,
                                               indices=loc.to(torch.int32), head_dim_v=self.kv_lora_rank)
        else:
            self.kv_buffer[layer.layer_id][loc] = cache_k.view(-1, 1, cache_k.shape[-1]).to(self.model_dtype)

    # ---- the formats either side of the path: PD-disaggregation transfer descriptors and the
    # radix cache's slot moves see the three tensors per layer, grouped by tensor kind
    def get_kv_size_bytes(self) -> int:
        """memory_pool.py:779-784."""
        total = 0
        for kv in self.kv_buffer:
            total += sum(t.nbytes for t in kv) if isinstance(kv, tuple) else kv.nbytes
        return total

    def get_contiguous_buf_infos(self):
        """memory_pool.py:787-800: (data_ptrs, data_lens, item_lens) of every transferable buffer; an item is one page.
        per_token_head: all layers' latent tensors, then all scale tensors, then all rope tensors."""
        if self.quant_method == "per_token_head":
            bufs = [kv[i] for i in range(3) for kv in self.kv_buffer]
        else:
            bufs = list(self.kv_buffer)
        return ([b.data_ptr() for b in bufs], [b.nbytes for b in bufs], [b[0].nbytes * self.page_size for b in bufs])

    def get_layerwise_buf_info_offsets(self, start_idx: int = 0):
        """memory_pool.py:802-806: for each layer, the indices of its buffers in get_contiguous_buf_infos()."""
        layer_num = len(self.kv_buffer)
        if self.quant_method == "per_token_head":
            return [[start_idx + i * layer_num + layer_id for i in range(3)] for layer_id in range(layer_num)]
        return [[start_idx + layer_id] for layer_id in range(layer_num)]

    def move_kv_cache(self, tgt_loc: torch.Tensor, src_loc: torch.Tensor):
        """memory_pool.py:747-764 (the native path): copy token slots src_loc -> tgt_loc in every layer, all three
        tensors -- codes, scales and pre-divided rope move together, so a moved token dequantises identically."""
        if tgt_loc.numel() == 0:
            return
        tgt, src = tgt_loc.view(-1).long(), src_loc.view(-1).long()
        for kv in self.kv_buffer:
            for buf in (kv if isinstance(kv, tuple) else (kv,)):
                buf[tgt] = buf[src]

    def get_key_split_contiguous(self, layer_id: int, indices: torch.Tensor):
        """memory_pool.py:817-824."""
        if self.quant_method == "per_token_head":
            k_lora, k_scale, k_rope = self.kv_buffer[layer_id]
            return flash_mla_fp8.dequantize_ckv_fused_indexed(k_lora, k_rope, k_scale, indices.to(torch.int32))
        rows = self.kv_buffer[layer_id][indices.long()]
        return rows[..., : self.kv_lora_rank].contiguous(), rows[..., self.kv_lora_rank:].contiguous()


class FlashMLADecodeMetadata:
    """flashmla_backend.py:32-46."""

    def __init__(self, flashmla_metadata=None, num_splits=None, block_table=None):
        self.flashmla_metadata = flashmla_metadata
        self.num_splits = num_splits
        self.block_table = block_table


class FlashMLABackend(AttentionBackend):
    """flashmla_backend.py:49-408, same constructor (:50-57).  Reads from ``model_runner``: model_config.{scaling,
    num_attention_heads, kv_lora_rank, qk_nope_head_dim, qk_rope_head_dim, v_head_dim, context_len},
    server_args.{speculative_algorithm, speculative_num_draft_tokens}, kv_cache_dtype, kv_cache_quant_method,
    kv_allocator, dtype; heads per rank = num_attention_heads // get_attention_tp_size() (:67-69)."""

    def __init__(self, model_runner, skip_prefill: bool = False, kv_indptr_buf=None, q_indptr_decode_buf=None,
                 multi_step_seqlen_incr: int = 0):
        cfg, sa = model_runner.model_config, model_runner.server_args
        attn_tp_size = get_attention_tp_size()
        # EXTEND / MIXED batches: the reference's base class (:59 super().__init__) -- here a delegate, when registered
        self.prefill_backend = None
        if _prefill_backend_factory is not None and not skip_prefill:
            self.prefill_backend = _prefill_backend_factory(model_runner, skip_prefill, kv_indptr_buf, q_indptr_decode_buf)
        self.softmax_scale = cfg.scaling
        self.draft_token_num = 0 if getattr(sa, "speculative_algorithm", None) is None else sa.speculative_num_draft_tokens
        self.multi_step_seqlen_incr = multi_step_seqlen_incr
        self.cache_dtype = model_runner.kv_cache_dtype
        self.cache_quant_method = model_runner.kv_cache_quant_method
        if self.cache_quant_method != "per_token_head" and self.cache_dtype == torch.float8_e4m3fn:
            # flashmla_backend.py:225-239 passes descale_q / descale_k that neither shipped kernel module accepts
            raise RuntimeError("FlashMLABackend (B200): an FP8 KV cache needs --kv-cache-quant-method per_token_head; "
                               "the reference's other FP8 branch is dead code (SURVEY.md 8a-2)")
        self.num_q_heads = cfg.num_attention_heads // attn_tp_size
        self.num_local_heads = self.num_q_heads
        self.kv_allocator = model_runner.kv_allocator
        self.forward_metadata: Optional[FlashMLADecodeMetadata] = None
        self.kv_lora_rank = cfg.kv_lora_rank
        self.qk_nope_head_dim = cfg.qk_nope_head_dim
        self.qk_rope_head_dim = cfg.qk_rope_head_dim
        self.v_head_dim = cfg.v_head_dim
        self.scaling = cfg.scaling
        self.q_data_type = model_runner.dtype
        self.kv_cache_dim = self.kv_lora_rank + self.qk_rope_head_dim
        self.max_context_len = cfg.context_len
        self.device = getattr(model_runner, "device", "cuda")

    def get_flash_mla_module(self, M):
        """flashmla_backend.py:18-22,84-85.  One module on B200: it serves every M (SwapAB below 64 q rows, CTA pairs
        above) for the per-token-FP8 pool and carries the bf16 kernel of both reference forks."""
        return flash_mla_fp8

    # ------------------------------------------------------------------ attention calls
    def _attend(self, q, layer, forward_batch, cache_seqlens):
        bs = forward_batch.batch_size
        if bs == 0:  # IDLE batch of a data-parallel attention rank: nothing to attend, an empty [0, h * 512] result
            return q.new_empty((0, layer.tp_q_head_num * layer.v_head_dim))
        k_cache = forward_batch.token_to_kv_pool.get_key_buffer(layer.layer_id)
        assert layer.tp_q_head_num == self.num_q_heads, f"{layer.tp_q_head_num=} != {self.num_q_heads=}"
        reshape_q = q.view(bs, -1, self.num_q_heads, layer.head_dim)
        mod = self.get_flash_mla_module(reshape_q.shape[1] * reshape_q.shape[2])
        if self.cache_quant_method != "per_token_head":
            # bf16 pool (flashmla_backend.py:240-256): the bf16 kernel, nothing quantised
            o, _ = mod.flash_mla_with_kvcache(
                q=reshape_q, k_cache=k_cache.view(-1, PAGE_SIZE, 1, self.kv_cache_dim),
                block_table=self.forward_metadata.block_table[:bs],
                cache_seqlens=cache_seqlens.to(torch.int32),
                head_dim_v=self.kv_lora_rank,
                tile_scheduler_metadata=self.forward_metadata.flashmla_metadata,
                num_splits=self.forward_metadata.num_splits,
                softmax_scale=layer.scaling, causal=True)
            return o.view(-1, layer.tp_q_head_num * layer.v_head_dim)
        q_nope, q_scale, q_rope = mod.quantize_ckv_per_token_head(reshape_q.contiguous(), self.kv_lora_rank)
        return self._attend_quantized(q_nope, q_scale, q_rope, layer, forward_batch, cache_seqlens)

    def _attend_quantized(self, q_nope, q_scale, q_rope, layer, forward_batch, cache_seqlens):
        """The attention call on an already quantised Q (flashmla_backend.py:207-224)."""
        bs = forward_batch.batch_size
        if bs == 0:
            return q_rope.new_empty((0, layer.tp_q_head_num * layer.v_head_dim))
        mod = self.get_flash_mla_module(q_nope.numel() // (bs * self.kv_lora_rank))
        k_cache_lora, k_scale, k_cache_rope = forward_batch.token_to_kv_pool.get_key_buffer(layer.layer_id)
        q_nope = q_nope.view(bs, -1, self.num_q_heads, self.kv_lora_rank)
        q_rope = q_rope.view(bs, -1, self.num_q_heads, self.qk_rope_head_dim)
        q_scale = q_scale.view(bs, -1, self.num_q_heads, 1)
        ms = getattr(self, "_merge_state", None)
        if ms is not None and ms.numel() < bs * ((q_nope.shape[1] * self.num_q_heads + 127) // 128) * 4:
            ms = None  # larger than the graph state was sized for: the module's per-stream default
        o, _ = mod.flash_mla_ckv_fp8_per_token(
            q_nope=q_nope, q_rope=q_rope,
            k_cache_lora=k_cache_lora.view(-1, PAGE_SIZE, 1, self.kv_lora_rank),
            k_cache_rope=k_cache_rope.view(-1, PAGE_SIZE, 1, self.qk_rope_head_dim),
            q_scale=q_scale, k_scale=k_scale.view(-1, PAGE_SIZE, 1, 1),
            block_table=self.forward_metadata.block_table[:bs],
            cache_seqlens=cache_seqlens.to(torch.int32),
            head_dim_v=self.kv_lora_rank,
            tile_scheduler_metadata=self.forward_metadata.flashmla_metadata,
            num_splits=self.forward_metadata.num_splits,
            softmax_scale=layer.scaling, causal=True, merge_state=ms)
        return o.view(-1, layer.tp_q_head_num * layer.v_head_dim)

    def forward_decode(self, q, k, v, layer, forward_batch, save_kv_cache=True, **kwargs):
        """flashmla_backend.py:178-256: write the new token's K (quantised) first, then attend over seq_lens.
        ``q`` may also be the (q_nope, q_scale, q_rope) triple of ``rope.fused_rope_quantize_set_kv`` -- the fused
        prologue has then rotated, quantised and cached already (call with k = v = None or save_kv_cache=False)."""
        if k is not None:
            assert v is not None
            if save_kv_cache:
                forward_batch.token_to_kv_pool.set_kv_buffer(layer, forward_batch.out_cache_loc, k, v)
        # flashmla_backend.py:202: seq_lens + multi_step_seqlen_incr (no elementwise launch per layer when the increment
        # is zero, i.e. outside the speculative draft loop)
        incr = self.multi_step_seqlen_incr
        cache_seqlens = forward_batch.seq_lens + incr if incr else forward_batch.seq_lens
        if isinstance(q, (tuple, list)):
            if self.cache_quant_method != "per_token_head":
                raise RuntimeError("a pre-quantised q needs the per_token_head KV pool")
            return self._attend_quantized(*q, layer, forward_batch, cache_seqlens)
        return self._attend(q, layer, forward_batch, cache_seqlens)

    def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache=True, q_pe=None, k_pe=None, **kwargs):
        """flashmla_backend.py:87-176: TARGET_VERIFY / DRAFT_EXTEND run the decode kernel with s_q = draft tokens."""
        mode = forward_batch.forward_mode
        if not (mode.is_target_verify() or mode.is_draft_extend()):
            # flashmla_backend.py:100-104: prefill of the base and draft models stays on the flashinfer MLA backend
            if self.prefill_backend is not None:
                return self.prefill_backend.forward_extend(q, k, v, layer, forward_batch, save_kv_cache, q_pe, k_pe)
            raise NotImplementedError("prefill (EXTEND) is outside the B200 decode hot path and no prefill backend is "
                                      "registered (flashmla_backend.set_prefill_backend_factory); the reference routes "
                                      "it to FlashInferMLAAttnBackend (flashmla_backend.py:100-104)")
        if k is not None:
            assert v is not None
            if save_kv_cache:
                forward_batch.token_to_kv_pool.set_kv_buffer(layer, forward_batch.out_cache_loc, k, v)
        cache_seqlens = forward_batch.seq_lens + self.draft_token_num
        if isinstance(q, (tuple, list)):  # the fused prologue's (q_nope, q_scale, q_rope), as in forward_decode
            if self.cache_quant_method != "per_token_head":
                raise RuntimeError("a pre-quantised q needs the per_token_head KV pool")
            return self._attend_quantized(*q, layer, forward_batch, cache_seqlens)
        return self._attend(q, layer, forward_batch, cache_seqlens)

    # ------------------------------------------------------------------ metadata
    def _metadata(self, seq_lens, forward_mode):
        if forward_mode.is_decode_or_idle():
            lens, m = seq_lens + self.multi_step_seqlen_incr, self.num_q_heads
        elif forward_mode.is_target_verify() or forward_mode.is_draft_extend():
            lens, m = seq_lens + self.draft_token_num, self.draft_token_num * self.num_q_heads
        else:
            raise RuntimeError(f"Not supported forward mode: {forward_mode}")
        # the scheduler's per-request cost tuned for this kernel (the reference's constant models its sm_90 kernel)
        mod = self.get_flash_mla_module(m)
        return mod.get_mla_metadata(lens.to(torch.int32).contiguous(), m, 1,
                                    fixed_overhead_blocks=mod.tuned_overhead_blocks(m),
                                    num_sm_parts=mod.tuned_num_sm_parts(m))

    def init_forward_metadata(self, forward_batch):
        """flashmla_backend.py:258-290 (other modes go to the base class, :289-290)."""
        mode = forward_batch.forward_mode
        if not (mode.is_decode_or_idle() or mode.is_target_verify() or mode.is_draft_extend()):
            if self.prefill_backend is not None:
                return self.prefill_backend.init_forward_metadata(forward_batch)
            raise RuntimeError(f"Not supported forward mode: {mode} (no prefill backend registered)")
        block_table = self.kv_allocator.req_to_page[forward_batch.req_pool_indices]
        mla_metadata, num_splits = self._metadata(forward_batch.seq_lens, forward_batch.forward_mode)
        self.forward_metadata = FlashMLADecodeMetadata(mla_metadata, num_splits, block_table)

    def init_cuda_graph_state(self, max_bs: int, block_kv_indices: Optional[torch.Tensor] = None):
        """flashmla_backend.py:292-326: static block table filled with page 1 (4 pages reserved for speculation) and
        static metadata sized for seq_len == 1 per padded row."""
        if block_kv_indices is None:
            max_context_len = self.max_context_len + PAGE_SIZE - 1
            block_kv_indices = torch.full((max_bs, (max_context_len + 4 * PAGE_SIZE) // PAGE_SIZE), 1,
                                          dtype=torch.int32, device=self.device)
        m = (self.draft_token_num or 1) * self.num_q_heads
        ones = torch.ones(max_bs, dtype=torch.int32, device=block_kv_indices.device)
        mod = self.get_flash_mla_module(m)
        self.cuda_graph_mla_metadata, self.cuda_graph_num_splits = mod.get_mla_metadata(
            ones, m, 1, fixed_overhead_blocks=mod.tuned_overhead_blocks(m), num_sm_parts=mod.tuned_num_sm_parts(m))
        self.cuda_graph_kv_indices = block_kv_indices
        # the graph's own zeroed merge-state scratch (flash_mla_ckv_fp8_per_token merges cut requests in the kernel)
        self._merge_state = flash_mla_fp8.new_merge_state(max_bs, 1, m, block_kv_indices.device)

    def init_forward_metadata_capture_cuda_graph(self, bs, num_tokens, req_pool_indices, seq_lens, forward_mode,
                                                 spec_info=None):
        """flashmla_backend.py:328-364."""
        block_table = self.kv_allocator.req_to_page[req_pool_indices]
        mla_metadata, num_splits = self._metadata(seq_lens, forward_mode)
        self.cuda_graph_mla_metadata.copy_(mla_metadata)
        self.cuda_graph_num_splits[: bs + 1].copy_(num_splits)
        self.cuda_graph_kv_indices[:bs, : block_table.shape[1]].copy_(block_table)
        self.forward_metadata = FlashMLADecodeMetadata(self.cuda_graph_mla_metadata,
                                                       self.cuda_graph_num_splits[: bs + 1],
                                                       self.cuda_graph_kv_indices[:bs, :])

    def init_forward_metadata_replay_cuda_graph(self, bs, req_pool_indices, seq_lens, seq_lens_sum, forward_mode,
                                                spec_info=None, seq_lens_cpu=None):
        """flashmla_backend.py:366-405: runs outside the graph; refreshes the static tensors in place."""
        req_pool_indices = req_pool_indices[:bs]
        block_table = self.kv_allocator.req_to_page[req_pool_indices]
        mla_metadata, num_splits = self._metadata(seq_lens[:bs], forward_mode)
        self.cuda_graph_mla_metadata.copy_(mla_metadata)
        self.cuda_graph_num_splits[: bs + 1].copy_(num_splits)
        self.cuda_graph_kv_indices[:bs, : block_table.shape[1]].copy_(block_table)
        self.forward_metadata.flashmla_metadata = self.cuda_graph_mla_metadata
        self.forward_metadata.num_splits = self.cuda_graph_num_splits[: bs + 1]
        self.forward_metadata.block_table = self.cuda_graph_kv_indices[:bs]

    def get_cuda_graph_seq_len_fill_value(self):
        return 1  # flashmla_backend.py:407-408


class FlashMLAMultiStepDecodeBackend:
    """flashmla_backend.py:411-477: one FlashMLABackend per speculative draft step, constructed as
    ``FlashMLAMultiStepDecodeBackend(draft_model_runner, speculative_num_steps)``
    (speculative/base_spec_worker.py:238-240).  Step i attends over ``seq_lens + i + 1`` (:418: the draft token the
    step itself just wrote is part of its context); metadata is prepared for the first ``spec_num_steps - 1`` steps
    (common_template, :424-433).

    The reference's ``init_cuda_graph_state`` / ``...capture_cuda_graph`` read ``self.speculative_num_steps`` and
    ``self.topk`` (:444,454), which its ``__init__`` never sets; both names are defined here (topk = 1, the only value
    the FlashMLA draft path supports) so those methods run."""

    def __init__(self, model_runner, spec_num_steps: int):
        self.spec_num_steps = self.speculative_num_steps = spec_num_steps
        self.topk = 1
        self.attn_backends = [FlashMLABackend(model_runner=model_runner, multi_step_seqlen_incr=i + 1, skip_prefill=True)
                              for i in range(spec_num_steps)]

    def common_template(self, forward_batch, call_fn):
        assert forward_batch.spec_info is not None
        for i in range(self.spec_num_steps - 1):
            call_fn(i, forward_batch)

    def init_forward_metadata(self, forward_batch):
        self.common_template(forward_batch, lambda i, fb: self.attn_backends[i].init_forward_metadata(fb))

    def init_cuda_graph_state(self, max_bs: int, max_num_tokens: Optional[int] = None):
        for b in self.attn_backends:
            b.init_cuda_graph_state(max_bs, block_kv_indices=None)

    def init_forward_metadata_capture_cuda_graph(self, forward_batch):
        def call_fn(i, fb):
            self.attn_backends[i].init_forward_metadata_capture_cuda_graph(
                fb.batch_size, fb.batch_size * self.topk, fb.req_pool_indices, fb.seq_lens,
                forward_mode=ForwardMode.DECODE, spec_info=fb.spec_info)

        self.common_template(forward_batch, call_fn)

    def init_forward_metadata_replay_cuda_graph(self, forward_batch, bs: int):
        def call_fn(i, fb):
            self.attn_backends[i].init_forward_metadata_replay_cuda_graph(
                bs, fb.req_pool_indices, fb.seq_lens, seq_lens_sum=-1, forward_mode=ForwardMode.DECODE,
                spec_info=fb.spec_info, seq_lens_cpu=None)

        self.common_template(forward_batch, call_fn)

    def get_cuda_graph_seq_len_fill_value(self):
        return 1
