"""CPU, world_size = 2 over gloo: the tensor-parallel partition of the path and its single all-reduce
(sglang_fluentllm_b200/tp.py).  The compute callbacks are the oracle's block-FP8 matmul and a dense softmax attention --
the checker standing in for the CUDA kernels, which is what lets the N > 1 host logic run without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from sglang_fluentllm_b200 import tp

HEADS, D_QK, D_V, Q_LORA, HIDDEN, B = 4, 192, 128, 256, 384, 5


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _quant_weight(w):
    """128x128 block quantisation (3rdparty/deep_gemm/tests/test_core.py:18-25) in numpy."""
    n, k = w.shape
    wb = w.reshape(n // 128, 128, k // 128, 128)
    amax = np.maximum(np.abs(wb).max(axis=(1, 3), keepdims=True), 1e-4)
    q = ofp8.f32_to_e4m3((wb * (448.0 / amax)).astype(np.float32).reshape(n, k))
    return torch.from_numpy(q), torch.from_numpy((amax / 448.0).reshape(n // 128, k // 128).astype(np.float32))


def _linear(x, ws):
    """act-quant 1x128 + block-FP8 matmul + bf16 rounding, all by the oracle."""
    w, s = ws
    xq, xs = ogemm.per_token_group_quant_fp8(x.float().numpy())
    y = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(xq), xs, ofp8.e4m3_to_f32(w.numpy()), s.numpy())
    return torch.from_numpy(ofp8.round_to_bf16(y.astype(np.float32)))


def _attention(q):  # [b, h, 192] -> [b, h, 128]: a fixed per-head map, enough to check that heads stay with their rank
    return torch.tanh(q[..., :D_V] + q[..., D_QK - D_V:].flip(-1))


def _inputs():
    rng = np.random.default_rng(0)
    w_q_b = _quant_weight(rng.standard_normal((HEADS * D_QK, Q_LORA)).astype(np.float32) / 16)
    w_o = _quant_weight(rng.standard_normal((HIDDEN, HEADS * D_V)).astype(np.float32) / 16)
    q_a = torch.from_numpy(ofp8.round_to_bf16(rng.standard_normal((B, Q_LORA)).astype(np.float32)))
    return w_q_b, w_o, q_a


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w_q_b, w_o, q_a = _inputs()
        plan = tp.TpPlan(rank, world, HEADS, D_QK, D_V)
        blk = tp.TensorParallelAttentionBlock(plan, w_q_b, w_o, _linear, _attention)
        y = blk.forward(q_a)
        ms = tp.max_over_ranks_ms(10.0 + rank)
        if rank == 0:
            out.put((y.numpy(), ms, plan.head_range))
    finally:
        dist.destroy_process_group()


def test_tp2_matches_unsharded_block():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    y, ms, head_range = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    w_q_b, w_o, q_a = _inputs()
    single = tp.TensorParallelAttentionBlock(tp.TpPlan(0, 1, HEADS, D_QK, D_V), w_q_b, w_o, _linear, _attention)
    ref = single.forward(q_a).numpy()
    # the ranks round their partial o_proj sums to bf16 before the all-reduce (as the real path does): 2 roundings
    assert np.abs(y - ref).max() <= 2.0 ** -7 * np.abs(ref).max() + 1e-6
    assert ms == 11.0            # max over ranks
    assert head_range == (0, 2)  # rank 0 of 2 owns heads 0-1


def test_plan_shards_and_alignment_rules():
    plan = tp.TpPlan(1, 2, 128, 192, 128)
    assert plan.heads_per_rank == 64 and plan.head_range == (64, 128)
    assert plan.column_slice() == slice(64 * 192, 128 * 192) and plan.row_slice() == slice(64 * 128, 128 * 128)
    w = torch.zeros(256, 384)
    s = torch.arange(2 * 3, dtype=torch.float32).reshape(2, 3)
    wc, sc = tp.shard_column_parallel(w, s, slice(128, 256))
    assert wc.shape == (128, 384) and torch.equal(sc, s[1:2])
    wr, sr = tp.shard_row_parallel(w, s, slice(128, 384))
    assert wr.shape == (256, 256) and torch.equal(sr, s[:, 1:3])
    with pytest.raises(ValueError, match="block_n"):
        tp.shard_column_parallel(w, s, slice(64, 192))
    with pytest.raises(ValueError, match="not divisible by tp size"):
        tp.TpPlan(0, 3, 128, 192, 128)
    with pytest.raises(ValueError, match="block_k"):   # 1 head x 64 wide per rank: not a multiple of 128
        tp.TpPlan(0, 2, 2, 128, 64)
    # tp = 1: the all-reduce is a no-op without a process group
    t = torch.ones(3)
    assert tp.all_reduce_sum(t) is t and tp.max_over_ranks_ms(3.5) == 3.5
