"""Exchanges either side of the attention block on one NVLink box.

Mirrors, for this path:
  * ``TritonRSAG`` (python/sglang/srt/distributed/device_communicators/custom_triton_rsag/triton_rsag.py:16-300): uneven
    reduce-scatter / all-gather over tokens inside a tensor-parallel group -- ``get_token_dist``, ``get_context``,
    ``reduce_scatter(hidden_states, tp_num_tokens | token_list_in_group) -> (rows of this rank, row offset)``,
    ``all_gather(...) -> all rows``;
  * the attention side of ``DecoderCommMananger`` (distributed/decoder_comm_manager.py:128-330): which exchange runs before
    and after the attention block for every combination of attention / dense / MoE parallel strategy;
  * ``tensor_model_parallel_all_reduce`` behind the row-parallel ``o_proj`` (layers/linear.py:1139).

CUDA tensors go over NVLink peer memory with this repo's own kernels (csrc/p2p_comm.cu through the C ABI:
``P2PGroup``); there is no other GPU path -- a missing library raises.  CPU tensors (the gloo tests of the host logic)
go through ``torch.distributed`` with uneven splits, which is also what the reference falls back to when its RSAG object
is absent (decoder_comm_manager.py:371-390).
"""
from __future__ import annotations

import ctypes
from enum import Enum
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib


class MoeParallelStrategy(str, Enum):     # distributed/parallel_strategy.py:4-6
    TENSOR_PARALLEL = "tp"
    EXPERT_PARALLEL = "ep"


class AttnParallelStrategy(str, Enum):    # :8-11
    TENSOR_PARALLEL = "tp"
    DATA_PARALLEL = "dp"
    COMBINE = "combine"


class DenseParallelStrategy(str, Enum):   # :13-16 (spelled DenseParallelStategy there)
    TENSOR_PARALLEL = "tp"
    REPLICATED = "rep"
    COMBINE = "combine"


DenseParallelStategy = DenseParallelStrategy  # the reference's spelling, for its call sites


def token_dist(total_tokens: int, world_size: int) -> List[int]:
    """triton_rsag.py:48-57: rows per rank, the first ``total % world`` ranks take one more."""
    return [total_tokens // world_size + (1 if r < total_tokens % world_size else 0) for r in range(world_size)]


class P2PGroup:
    """This rank's communication buffer and flag pad, its peers' mapped into this process, and the collective kernels.

    ``group``: the ``torch.distributed`` group whose ranks live on one node with peer access (it is only used to exchange
    the 64-byte IPC handles once).  ``max_bytes``: the largest operand of a call; the buffer is twice that (its halves
    alternate per call, see csrc/p2p_comm.cu)."""

    def __init__(self, group, device: torch.device, max_bytes: int):
        self.group = group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = torch.device(device)
        self.max_operand_bytes = (int(max_bytes) + 31) // 32 * 32
        self.max_bytes = 2 * self.max_operand_bytes
        L = _lib.lib()
        self._L = L
        with torch.cuda.device(self.device):
            buf, flg = ctypes.c_void_p(), ctypes.c_void_p()
            _lib.check(L.b200_comm_alloc(self.max_bytes, ctypes.byref(buf)), "comm_alloc")
            _lib.check(L.b200_comm_alloc(L.b200_comm_flag_bytes(), ctypes.byref(flg)), "comm_alloc")
            self._own = (buf.value, flg.value)
            handles = []
            for p in self._own:
                h = ctypes.create_string_buffer(64)
                _lib.check(L.b200_comm_ipc_export(p, h), "comm_ipc_export")
                handles.append(h.raw)
            gathered: List[Optional[Tuple[bytes, bytes]]] = [None] * self.world
            dist.all_gather_object(gathered, tuple(handles), group=group)
            self._opened: List[int] = []
            bufs, flags = [], []
            for r, hs in enumerate(gathered):
                if r == self.rank:
                    bufs.append(self._own[0])
                    flags.append(self._own[1])
                    continue
                ptrs = []
                for h in hs:
                    q = ctypes.c_void_p()
                    _lib.check(L.b200_comm_ipc_open(ctypes.create_string_buffer(h, 64), ctypes.byref(q)), "comm_ipc_open")
                    ptrs.append(q.value)
                    self._opened.append(q.value)
                bufs.append(ptrs[0])
                flags.append(ptrs[1])
            self._bufs = (ctypes.c_void_p * self.world)(*bufs)
            self._flags = (ctypes.c_void_p * self.world)(*flags)
        dist.barrier(group=group)  # every peer has mapped every buffer before anybody's first kernel polls a flag

    def _offsets(self, token_list: Sequence[int]):
        if len(token_list) != self.world:
            raise RuntimeError("token list has %d entries for %d ranks" % (len(token_list), self.world))
        off = [0]
        for t in token_list:
            off.append(off[-1] + int(t))
        return off, (ctypes.c_int64 * (self.world + 1))(*off)

    def enable_nvls(self) -> bool:
        """Collective.  Map a symmetric buffer with a multicast address over the group (NVSwitch in-switch reduction,
        ``multimem.ld_reduce`` / ``multimem.st``): torch's symmetric-memory allocator does the driver work (the reference
        gets its multicast pointer the same way, triton_rsag.py:38-46,190).  Returns whether the box supports it; the
        reason it does not is kept in ``nvls_error``."""
        self._nvls = None
        self.nvls_error = None
        try:
            import torch.distributed._symmetric_memory as symm_mem
            with torch.cuda.device(self.device):
                t = symm_mem.empty(self.max_bytes, dtype=torch.uint8, device=self.device)
                hdl = symm_mem.rendezvous(t, self.group)
                mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                if mc == 0:
                    self.nvls_error = "the symmetric-memory handle has no multicast address (no NVSwitch multicast here)"
                else:
                    self._nvls = (t, hdl, mc)
        except Exception as e:  # allocator / driver without fabric or multicast support
            self.nvls_error = "%s: %s" % (type(e).__name__, e)
        ok = torch.tensor([1 if self._nvls is not None else 0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)  # all ranks or none
        if int(ok.item()) == 0:
            self._nvls = None
        return self._nvls is not None

    def all_reduce(self, t: torch.Tensor, out: Optional[torch.Tensor] = None, nvls: Optional[bool] = None) -> torch.Tensor:
        """Sum of every rank's ``t`` into ``out`` (default: in place in ``t``).  ``nvls``: through the switch (needs
        enable_nvls()); default: whenever it is enabled."""
        self._check(t)
        out = t if out is None else out
        # measured (tests/test_gpu_tp2.py, 1.8 MB): 2 GPUs 23 us peer loads / 28 us switch / 29 us NCCL; 8 GPUs 36 / 29 / 32
        use_nvls = (getattr(self, "_nvls", None) is not None and self.world > 2) if nvls is None else nvls
        if use_nvls:
            if getattr(self, "_nvls", None) is None:
                raise RuntimeError("P2PGroup.all_reduce(nvls=True): enable_nvls() has not succeeded (%s)" % getattr(self, "nvls_error", None))
            buf, _, mc = self._nvls
            _lib.check(self._L.b200_comm_all_reduce_nvls_bf16(mc, buf.data_ptr(), self._flags, self.max_bytes, self.rank,
                                                              self.world, t.data_ptr(), t.numel(), out.data_ptr(), 0,
                                                              _lib.current_stream()), "comm_all_reduce_nvls")
            return out
        _lib.check(self._L.b200_comm_all_reduce_bf16(self._bufs, self._flags, self.max_bytes, self.rank, self.world,
                                                     t.data_ptr(), t.numel(), out.data_ptr(), 0, _lib.current_stream()),
                   "comm_all_reduce")
        return out

    def reduce_scatter(self, t: torch.Tensor, token_list: Sequence[int]) -> Tuple[torch.Tensor, int]:
        """``t`` [sum(token_list), hidden] on every rank -> (the sum's rows of this rank, their row offset)."""
        self._check(t)
        off, c_off = self._offsets(token_list)
        if t.shape[0] != off[-1]:
            raise RuntimeError("reduce_scatter: %d rows for a token list summing to %d" % (t.shape[0], off[-1]))
        out = torch.empty(off[self.rank + 1] - off[self.rank], t.shape[1], dtype=t.dtype, device=t.device)
        _lib.check(self._L.b200_comm_reduce_scatter_bf16(self._bufs, self._flags, self.max_bytes, self.rank, self.world,
                                                         c_off, t.data_ptr(), t.shape[1], out.data_ptr(), 0,
                                                         _lib.current_stream()), "comm_reduce_scatter")
        return out, off[self.rank]

    def all_gather(self, t: torch.Tensor, token_list: Sequence[int]) -> torch.Tensor:
        """``t`` [token_list[rank], hidden] -> [sum(token_list), hidden], every rank's rows in rank order."""
        self._check(t)
        off, c_off = self._offsets(token_list)
        if t.shape[0] != off[self.rank + 1] - off[self.rank]:
            raise RuntimeError("all_gather: %d rows where the token list gives this rank %d" % (
                t.shape[0], off[self.rank + 1] - off[self.rank]))
        out = torch.empty(off[-1], t.shape[1], dtype=t.dtype, device=t.device)
        _lib.check(self._L.b200_comm_all_gather_bf16(self._bufs, self._flags, self.max_bytes, self.rank, self.world, c_off,
                                                     t.data_ptr(), t.shape[1], out.data_ptr(), 0, _lib.current_stream()),
                   "comm_all_gather")
        return out

    def _check(self, t: torch.Tensor) -> None:
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous()):
            raise RuntimeError("P2PGroup: contiguous bfloat16 CUDA tensors only")  # triton_rsag.py:244: bf16 only there too
        if t.numel() % 8 or (t.dim() == 2 and t.shape[1] % 8):
            raise RuntimeError("P2PGroup: rows must be multiples of 16 bytes")
        if t.numel() * 2 > self.max_operand_bytes:
            raise RuntimeError("P2PGroup: %d bytes exceed the %d this group was sized for" % (t.numel() * 2, self.max_operand_bytes))

    def close(self) -> None:
        for p in self._opened:
            self._L.b200_comm_ipc_close(p)
        self._opened = []
        for p in self._own:
            self._L.b200_comm_free(p)
        self._own = ()


class RSAG:
    """``TritonRSAG`` for this box (triton_rsag.py:16-300): same methods, same results."""

    def __init__(self, group, rank_in_group: int, max_tokens: int, hidden_size: int, device=None):
        self.group = group
        self.rank_in_group = rank_in_group
        self.world_size = dist.get_world_size(group)
        self.max_tokens = max_tokens
        self.hidden_size = hidden_size
        self.dtype = torch.bfloat16
        self.device = torch.device(device) if device is not None else torch.device("cpu")
        self.p2p = (P2PGroup(group, self.device, max_tokens * hidden_size * 2) if self.device.type == "cuda" else None)

    def get_token_dist(self, total_tokens_in_group: int) -> List[int]:
        return token_dist(total_tokens_in_group, self.world_size)

    def get_context(self, token_list_in_group: Sequence[int]) -> Tuple[int, int, int]:
        """(total rows, this rank's rows, this rank's first row) -- triton_rsag.py:59-71."""
        total = sum(token_list_in_group)
        assert total <= self.max_tokens, f"The inner comm buffer is too small: {total=} is not <= {self.max_tokens=}"
        return total, token_list_in_group[self.rank_in_group], sum(token_list_in_group[: self.rank_in_group])

    def _tokens(self, tp_num_tokens, token_list_in_group):
        assert tp_num_tokens is not None or token_list_in_group is not None, \
            "Either tp_num_tokens or token_list_in_group must be provided"
        return list(token_list_in_group) if token_list_in_group is not None else self.get_token_dist(tp_num_tokens)

    def reduce_scatter(self, hidden_states: torch.Tensor, tp_num_tokens: Optional[int] = None,
                       token_list_in_group: Optional[Sequence[int]] = None, safe: bool = True) -> Tuple[torch.Tensor, int]:
        tokens = self._tokens(tp_num_tokens, token_list_in_group)
        assert hidden_states.dtype == torch.bfloat16, "Only bfloat16 is supported for now"
        total, local, offset = self.get_context(tokens)
        assert hidden_states.shape[0] == total and hidden_states.shape[-1] == self.hidden_size, \
            f"Dismatched shape, {hidden_states.shape[0]=} != {total=} or {hidden_states.shape[-1]=} != {self.hidden_size=}"
        if hidden_states.is_cuda:
            return self.p2p.reduce_scatter(hidden_states.contiguous(), tokens)
        out = torch.empty(local, hidden_states.shape[-1], dtype=hidden_states.dtype)
        # gloo has no reduce_scatter: sum everything, keep this rank's rows (host-logic tests only)
        full = hidden_states.float().clone()
        dist.all_reduce(full, group=self.group)
        out.copy_(full[offset: offset + local])
        return out, offset

    def all_gather(self, hidden_states: torch.Tensor, tp_num_tokens: Optional[int] = None,
                   token_list_in_group: Optional[Sequence[int]] = None, safe: bool = True) -> torch.Tensor:
        tokens = self._tokens(tp_num_tokens, token_list_in_group)
        assert hidden_states.dtype == torch.bfloat16, "Only bfloat16 is supported for now"
        total, local, offset = self.get_context(tokens)
        assert hidden_states.shape[0] == local and hidden_states.shape[-1] <= self.hidden_size, \
            f"{hidden_states.shape=}|{local=}|{hidden_states.device=} Dismatched shape"
        if hidden_states.is_cuda:
            return self.p2p.all_gather(hidden_states.contiguous(), tokens)
        # gloo gathers equal shapes only: pad to the largest share, trim afterwards (host-logic tests only)
        width, most = hidden_states.shape[-1], max(tokens)
        padded = torch.zeros(most, width, dtype=hidden_states.dtype)
        padded[:local] = hidden_states
        parts = [torch.empty(most, width, dtype=hidden_states.dtype) for _ in tokens]
        dist.all_gather(parts, padded, group=self.group)
        return torch.cat([p[:n] for p, n in zip(parts, tokens)])


def _unchanged(h, residual, *_):
    return h, residual


class DecoderCommManager:
    """The attention side of decoder_comm_manager.py:128-330 as a table: for every (layer kind, attention strategy,
    dense / MoE strategy) the exchange in front of the attention block, the one behind it, and the one behind the final
    norm.  ``all_reduce(t) -> t`` and the RSAG objects are injected (``attn_rsag`` may be None when attn_tp_size == 1)."""

    def __init__(self, layer_id: int, attn_parallel_strategy, dense_parallel_strategy, moe_parallel_strategy,
                 is_moe_layer: bool, num_layers: int, *, attn_tp_size: int, dense_tp_size: int, tp_size: int,
                 attn_rsag: Optional[RSAG] = None, all_reduce: Optional[Callable] = None):
        self.layer_id, self.num_layers, self.is_moe_layer = layer_id, num_layers, is_moe_layer
        self.attn = AttnParallelStrategy(attn_parallel_strategy)
        self.dense = DenseParallelStrategy(dense_parallel_strategy)
        self.moe = MoeParallelStrategy(moe_parallel_strategy)
        self.attn_tp_size, self.dense_tp_size, self.tp_size = attn_tp_size, dense_tp_size, tp_size
        self.attn_rsag = attn_rsag if attn_tp_size > 1 else None
        self.all_reduce = all_reduce

    # which exchange follows the attention block: "ar" all-reduce, "rs" reduce-scatter over tokens, None nothing
    def _post_attn_kind(self) -> Optional[str]:
        A, D, M = AttnParallelStrategy, DenseParallelStrategy, MoeParallelStrategy
        if self.attn == A.DATA_PARALLEL:
            if self.is_moe_layer or self.dense in (D.REPLICATED, D.TENSOR_PARALLEL):
                return None
        elif self.attn == A.TENSOR_PARALLEL:
            if self.is_moe_layer:
                return "ar" if self.moe == M.TENSOR_PARALLEL else "rs"
            if self.dense == D.REPLICATED:
                return "rs"
            if self.dense == D.TENSOR_PARALLEL:
                uneven = self.moe == M.EXPERT_PARALLEL or self.attn_tp_size != self.tp_size
                return "rs" if uneven else "ar"
        raise NotImplementedError("NotImplementedError")  # decoder_comm_manager.py:217,255

    def _gathers_before_attn(self) -> bool:
        A, D, M = AttnParallelStrategy, DenseParallelStrategy, MoeParallelStrategy
        if self.attn != A.TENSOR_PARALLEL:
            return False
        if self.is_moe_layer:
            return self.moe == M.EXPERT_PARALLEL
        if self.dense == D.REPLICATED:
            return True
        if self.dense == D.TENSOR_PARALLEL:
            return self.moe == M.EXPERT_PARALLEL or self.attn_tp_size != self.tp_size
        return False

    def pre_attn_comm(self, hidden_states, tp_num_tokens, is_second_attn: bool = False):
        """:152-184: from the second attention on, the rows scattered behind the previous block come back together."""
        if (self.layer_id > 0 or is_second_attn) and self._gathers_before_attn():
            return self.attn_rsag.all_gather(hidden_states, tp_num_tokens)
        return hidden_states

    def post_attn_comm(self, hidden_states, residual, tp_num_tokens, is_first_attn: bool = True):
        """:186-262.  After a reduce-scatter in front of the very first residual add, the residual is cut to this rank's
        rows as well."""
        kind = self._post_attn_kind()
        if kind is None:
            return hidden_states, residual
        if kind == "ar":
            return self.all_reduce(hidden_states), residual
        hidden_states, offset = self.attn_rsag.reduce_scatter(hidden_states, tp_num_tokens)
        if self.layer_id == 0 and is_first_attn:
            residual = residual[offset: offset + hidden_states.shape[0]]
        return hidden_states, residual

    def post_final_norm_comm(self, hidden_states, residual, tp_num_tokens):
        """:296-312: the logits head wants every row on every rank again."""
        if self.attn == AttnParallelStrategy.TENSOR_PARALLEL:
            if self.moe == MoeParallelStrategy.EXPERT_PARALLEL or self.dense_tp_size != self.tp_size:
                return self.attn_rsag.all_gather(hidden_states, tp_num_tokens), residual
        return hidden_states, residual


DecoderCommMananger = DecoderCommManager  # the reference's spelling
