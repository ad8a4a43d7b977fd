"""One-box tensor parallelism of the decode hot path: attention heads and column-parallel
weights are split over the ranks with no exchange, the row-parallel ``o_proj`` output meets in ONE all-reduce.

Reference behaviour mirrored here:
  * heads per rank = num_heads / attn_tp (python/sglang/srt/layers/attention/flashmla_backend.py:67-75); MLA has a
    single latent KV head, so every rank holds and streams the full KV pool
    (model_executor/attn_initializer.py:261-275) -- only q / out are sharded;
  * column-parallel (q_b_proj, kv_b_proj) splits N, row-parallel (o_proj) splits K, both on 128-block boundaries of the
    block-quantised weight (python/sglang/srt/layers/dense/layouts/fp8.py:118-142, layers/linear.py:199,963);
  * the row-parallel output is summed with one all-reduce (layers/linear.py:1139).

The collective is ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests); the compute callbacks
are injected, so the sharding / reduction logic is testable on CPU with the oracle as the compute (tests/test_tp_gloo.py)
while bench.py passes the CUDA kernels.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist

BLOCK = 128  # weight quantisation block (both dimensions)


@dataclass(frozen=True)
class TpPlan:
    """Which slice of the attention block a rank owns."""
    rank: int
    size: int
    num_heads: int
    qk_head_dim: int   # per-head width of q_b_proj's output (nope + rope), 192 for DeepSeek-V3
    v_head_dim: int    # per-head width of o_proj's input, 128 for DeepSeek-V3

    def __post_init__(self):
        if self.size < 1 or not (0 <= self.rank < self.size):
            raise ValueError(f"bad rank {self.rank} of {self.size}")
        if self.num_heads % self.size != 0:
            raise ValueError(f"num_heads = {self.num_heads} is not divisible by tp size = {self.size}")
        if (self.heads_per_rank * self.qk_head_dim) % BLOCK != 0:
            raise ValueError(f"Weight output_partition_size = {self.heads_per_rank * self.qk_head_dim} is not divisible by "
                             f"weight quantization block_n = {BLOCK}.")
        if (self.heads_per_rank * self.v_head_dim) % BLOCK != 0:
            raise ValueError(f"Weight input_size_per_partition = {self.heads_per_rank * self.v_head_dim} is not divisible "
                             f"by weight quantization block_k = {BLOCK}.")

    @property
    def heads_per_rank(self) -> int:
        return self.num_heads // self.size

    @property
    def head_range(self) -> Tuple[int, int]:
        h = self.heads_per_rank
        return self.rank * h, (self.rank + 1) * h

    def column_slice(self) -> slice:
        """Rows of q_b_proj's weight [num_heads * qk_head_dim, q_lora] this rank owns."""
        h0, h1 = self.head_range
        return slice(h0 * self.qk_head_dim, h1 * self.qk_head_dim)

    def row_slice(self) -> slice:
        """Columns of o_proj's weight [hidden, num_heads * v_head_dim] this rank owns."""
        h0, h1 = self.head_range
        return slice(h0 * self.v_head_dim, h1 * self.v_head_dim)


def shard_column_parallel(weight: torch.Tensor, scale: torch.Tensor, rows: slice):
    """(weight [n, k], scale [n/128, k/128]) -> the rank's N slice; ``rows`` must sit on 128-row block boundaries."""
    if rows.start % BLOCK or rows.stop % BLOCK:
        raise ValueError(f"column-parallel slice {rows.start}:{rows.stop} is not aligned to block_n = {BLOCK}")
    return weight[rows].contiguous(), scale[rows.start // BLOCK: rows.stop // BLOCK].contiguous()


def shard_row_parallel(weight: torch.Tensor, scale: torch.Tensor, cols: slice):
    """(weight [n, k], scale [n/128, k/128]) -> the rank's K slice; ``cols`` must sit on 128-column block boundaries."""
    if cols.start % BLOCK or cols.stop % BLOCK:
        raise ValueError(f"row-parallel slice {cols.start}:{cols.stop} is not aligned to block_k = {BLOCK}")
    return weight[:, cols].contiguous(), scale[:, cols.start // BLOCK: cols.stop // BLOCK].contiguous()


_p2p = None  # comm.P2PGroup of the tensor-parallel group once enable_p2p_all_reduce() has run


def enable_p2p_all_reduce(group, device, max_bytes: int):
    """Route the block's all-reduce through this repo's NVLink peer-memory kernel (csrc/p2p_comm.cu) instead of NCCL:
    collective (every rank of ``group`` calls it once, before any graph capture).  Returns the comm.P2PGroup."""
    global _p2p
    from . import comm
    _p2p = comm.P2PGroup(group if group is not None else dist.group.WORLD, device, max_bytes)
    if _p2p.world > 2:
        _p2p.enable_nvls()  # in-switch reduction where the box has it (falls back to peer loads, reason in nvls_error)
    return _p2p


def disable_p2p_all_reduce() -> None:
    global _p2p
    if _p2p is not None:
        _p2p.close()
    _p2p = None


def all_reduce_sum(t: torch.Tensor, group=None) -> torch.Tensor:
    """The path's single exchange.  In place; a no-op without a process group (tp = 1).  CUDA bf16 tensors that fit the
    peer-memory buffer go through the P2P kernel when it is enabled, everything else through torch.distributed (NCCL on
    the GPU box, gloo in the CPU tests)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        if (_p2p is not None and t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.numel() % 8 == 0
                and t.numel() * 2 <= _p2p.max_operand_bytes):
            _p2p.all_reduce(t)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


class TensorParallelAttentionBlock:
    """q_b_proj (column-parallel) -> attention over this rank's heads -> o_proj (row-parallel) -> all-reduce.

    ``linear(x, (w, s)) -> y`` and ``attention(q [b, h_local, d_qk]) -> o [b, h_local, d_v]`` are injected."""

    def __init__(self, plan: TpPlan, w_q_b, w_o, linear: Callable, attention: Callable, group=None):
        self.plan = plan
        self.w_q_b = shard_column_parallel(w_q_b[0], w_q_b[1], plan.column_slice())
        self.w_o = shard_row_parallel(w_o[0], w_o[1], plan.row_slice())
        self.linear, self.attention, self.group = linear, attention, group

    def forward(self, q_a: torch.Tensor) -> torch.Tensor:
        p = self.plan
        q = self.linear(q_a, self.w_q_b)                                    # [b, h_local * d_qk], no exchange
        o = self.attention(q.view(q.shape[0], p.heads_per_rank, p.qk_head_dim))  # [b, h_local, d_v], no exchange
        y = self.linear(o.reshape(o.shape[0], p.heads_per_rank * p.v_head_dim), self.w_o)  # partial sums over K
        return all_reduce_sum(y, self.group)


def max_over_ranks_ms(ms: float, device: Optional[torch.device] = None, group=None) -> float:
    """Timing rule of bench.py: a multi-GPU number is the maximum over the ranks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
