"""CPU baseline: port of the reference's torch-native CPU path for the benchmarked step (TEST / BENCH INFRASTRUCTURE
ONLY -- never on the product path).

  * decode attention: TorchNativeAttnBackend._run_sdpa_forward_decode
    (python/sglang/srt/layers/attention/torch_native_backend.py:275-343): per-sequence Python loop, gather
    k_cache[req_to_token[req, :len]], torch scaled_dot_product_attention with enable_gqa (absorbed MLA: q 576 wide,
    v = first 512 columns of the same bf16 pool, torch_native_backend.py:472-528, memory_pool.py:841-849).  The
    torch-native backend cannot read the per-token-FP8 tuple pool, so it runs on a bf16 pool.
  * block-FP8 linear: native_w8a8_block_fp8_matmul (python/sglang/test/test_block_fp8.py:89-141) with
    native_per_token_group_quant_fp8 (:15-40).
The reference package itself is not importable without stubs, and /root/reference does not exist on
the GPU box, so there this port is what is timed ("kind": "port" in bench.py); where the tree exists, bench.py times the
reference's own functions through oracle/ref_import.py ("kind": "reference") with the same driver below.
"""
from __future__ import annotations

import time

import torch
from torch.nn.functional import scaled_dot_product_attention


def run_sdpa_forward_decode(query, output, k_cache, v_cache, req_to_token, req_pool_indices, seq_lens, scaling=None,
                            enable_gqa=False, causal=False):
    """torch_native_backend.py:275-343 (same loop structure and tensor movement)."""
    query = query.movedim(0, query.dim() - 2)
    start_q, start_kv = 0, 0
    for seq_idx in range(seq_lens.shape[0]):
        seq_len_q = 1
        seq_len_kv = seq_lens[seq_idx]
        end_q = start_q + seq_len_q
        end_kv = start_kv + seq_len_kv
        per_req_query = query[:, start_q:end_q, :]
        req_pool_idx = req_pool_indices[seq_idx]
        per_req_tokens = req_to_token[req_pool_idx, :seq_len_kv]
        per_req_key = k_cache[per_req_tokens].movedim(0, query.dim() - 2)
        per_req_value = v_cache[per_req_tokens].movedim(0, query.dim() - 2)
        per_req_out = (
            scaled_dot_product_attention(per_req_query.unsqueeze(0), per_req_key.unsqueeze(0),
                                         per_req_value.unsqueeze(0), enable_gqa=enable_gqa, scale=scaling,
                                         is_causal=causal).squeeze(0).movedim(query.dim() - 2, 0))
        output[start_q:end_q, :, :] = per_req_out
        start_q, start_kv = end_q, end_kv
    return output


def native_per_token_group_quant_fp8(x, group_size=128, eps=1e-10, dtype=torch.float8_e4m3fn):
    """test_block_fp8.py:15-40."""
    finfo = torch.finfo(dtype)
    x_ = x.reshape(x.numel() // group_size, group_size)
    amax = x_.abs().max(dim=-1, keepdim=True)[0].clamp(min=eps).to(torch.float32)
    x_s = amax / finfo.max
    x_q = (x_ / x_s).clamp(min=finfo.min, max=finfo.max).to(dtype)
    return x_q.reshape(x.shape), x_s.reshape(x.shape[:-1] + (x.shape[-1] // group_size,))


def native_w8a8_block_fp8_matmul(A, B, As, Bs, block_size=(128, 128), output_dtype=torch.bfloat16):
    """test_block_fp8.py:89-141."""
    A = A.to(torch.float32)
    B = B.to(torch.float32)
    block_n, block_k = block_size
    M = A.numel() // A.shape[-1]
    N, K = B.shape
    A = A.reshape(M, A.shape[-1])
    As = As.reshape(M, As.shape[-1])
    n_tiles = (N + block_n - 1) // block_n
    k_tiles = (K + block_k - 1) // block_k
    C = torch.zeros((M, N), dtype=torch.float32)
    for i in range(k_tiles):
        a = A[:, i * block_k:min((i + 1) * block_k, K)]
        for j in range(n_tiles):
            b = B[j * block_n:min((j + 1) * block_n, N), i * block_k:min((i + 1) * block_k, K)]
            s = As[:, i:i + 1] * Bs[j][i]
            C[:, j * block_n:min((j + 1) * block_n, N)] += torch.matmul(a, b.t()) * s
    return C.to(output_dtype)


def time_decode_attention(b, seq, heads, repeats=2, decode=None):
    """Seconds for one layer of decode attention over `b` requests (median of `repeats` after one warm-up).  `decode`:
    the loop to time (default: the port above; bench.py passes the reference's own method where it can import it)."""
    decode = decode or run_sdpa_forward_decode
    g = torch.Generator().manual_seed(0)
    total = b * seq
    k_cache = torch.randn(total + 64, 1, 576, dtype=torch.bfloat16, generator=g)
    v_cache = k_cache[..., :512]
    req_to_token = (torch.randperm(total, generator=g) + 64).view(b, seq)
    q = torch.randn(b, heads, 576, dtype=torch.bfloat16, generator=g)
    out = torch.empty(b, heads, 512, dtype=torch.bfloat16)
    seq_lens = torch.full((b,), seq, dtype=torch.int64)
    pool_idx = torch.arange(b)
    ts = []
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        decode(q, out, k_cache, v_cache, req_to_token, pool_idx, seq_lens, scaling=576 ** -0.5, enable_gqa=True)
        ts.append(time.perf_counter() - t0)
    return sorted(ts[1:])[len(ts[1:]) // 2]


def time_block_fp8_linears(m, heads, repeats=1, matmul=None, quant=None):
    """Seconds for the three block-FP8 linears of one attention block (qkv_a, q_b, o_proj shapes) on `m` rows.
    `matmul` / `quant`: the functions to time (default: the ports above)."""
    matmul = matmul or native_w8a8_block_fp8_matmul
    quant = quant or native_per_token_group_quant_fp8
    g = torch.Generator().manual_seed(1)
    shapes = [(1536 + 576, 7168), (heads * 192, 1536), (7168, heads * 128)]
    ws = []
    for n, k in shapes:
        w = torch.randn(n, k, generator=g) * k ** -0.5
        wq, w_s = native_per_token_group_quant_fp8(w.reshape(-1, 128).reshape(n, k), 128)
        ws.append((wq, torch.ones((n + 127) // 128, k // 128)))
    ts = []
    for i in range(repeats + 1):
        t0 = time.perf_counter()
        for (n, k), (wq, w_s) in zip(shapes, ws):
            x = torch.randn(m, k, dtype=torch.bfloat16, generator=g)
            xq, xs = quant(x, 128)
            matmul(xq, wq, xs, w_s, [128, 128], torch.bfloat16)
        ts.append(time.perf_counter() - t0)
    return sorted(ts[1:])[len(ts[1:]) // 2]
