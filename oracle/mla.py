"""Oracle: per-token-FP8 paged MLA decode path (TEST INFRASTRUCTURE ONLY).

Restates, on the CPU with numpy, the arithmetic of the reference's
``flash_mla_fp8`` extension (3rdparty/flashmla-fp8, pinned a29b228).  All
reference paths below are relative to /root/reference/3rdparty/flashmla-fp8/.

The reference kernels are sm_90a-only (wgmma; ``TORCH_CHECK(is_sm90)`` at
csrc/flash_api.cpp:233-235) and are built with ``--use_fast_math``
(setup.py:69), so no bit-level golden output of them can be produced on this
box.  This restatement uses IEEE float32 (div.rn, no ftz) where the reference
would use the approximate forms; the product CUDA kernels do the same, so the
byte/integer outputs (e4m3 codes, scale bits, scheduler metadata) are compared
bit-exactly against this file.  PARITY UNPINNED for the attention kernel itself
(see oracle/__init__.py).
"""
from __future__ import annotations

import math

import numpy as np

from .fp8 import FP8_MAX, e4m3_to_f32, f32_to_e4m3, round_to_bf16

PAGE_BLOCK_SIZE = 64          # csrc/kernels/ckv_fp8_per_token/config.h:6
BLOCK_SIZE_M = 64             # config.h:5
FIXED_OVERHEAD_NUM_BLOCKS = 5 # config.h:14
TILE_SCHEDULER_METADATA_SIZE = 8  # csrc/kernels/ckv_fp8_per_token/params.h:72
HEAD_DIM_V = 512
HEAD_DIM_ROPE = 64

MAX_INIT_VAL_SM = np.float32(-1e30)   # splitkv_mla.cu:16
MAX_INIT_VAL = np.float32(-1e33)      # splitkv_mla.cu:17
MAX_INIT_VAL_P = np.float32(1e-26)    # splitkv_mla.cu:18
LOG2E = np.float32(math.log2(math.e))


def ceil_div(a: int, b: int) -> int:
    return (a + b - 1) // b


# --------------------------------------------------------------------------- #
# a-4  get_mla_metadata                                                        #
# --------------------------------------------------------------------------- #
def num_sm_parts_for(num_heads_per_head_k: int, num_heads_k: int, sm_count: int) -> int:
    """csrc/flash_api.cpp:41-43."""
    return sm_count // num_heads_k // ceil_div(num_heads_per_head_k, BLOCK_SIZE_M)


def get_mla_metadata(seqlens_k, num_heads_per_head_k: int, num_heads_k: int, sm_count: int,
                     fixed_overhead_blocks: int = FIXED_OVERHEAD_NUM_BLOCKS):
    """Integer tile scheduler.  csrc/kernels/get_mla_metadata.cu:8-75 (kernel body),
    csrc/flash_api.cpp:27-63 (num_sm_parts, output shapes).

    Returns (tile_scheduler_metadata int32 [num_sm_parts, 8], num_splits int32 [b+1]).
    Columns 5..7 of the metadata are never written by the reference (torch.empty);
    the oracle zeroes them and tests compare columns 0..4 only.
    """
    seqlens_k = np.asarray(seqlens_k, dtype=np.int64)
    batch_size = int(seqlens_k.shape[0])
    block_size_n = PAGE_BLOCK_SIZE
    fixed = int(fixed_overhead_blocks)  # the reference's constant unless a test sweeps the product's extension
    num_sm_parts = num_sm_parts_for(num_heads_per_head_k, num_heads_k, sm_count)

    num_blocks = [ceil_div(int(s), block_size_n) for s in seqlens_k]
    total_num_blocks = sum(nb + fixed for nb in num_blocks)
    payload = max(ceil_div(total_num_blocks, num_sm_parts) + fixed, 2 * fixed)

    meta = np.zeros((num_sm_parts, TILE_SCHEDULER_METADATA_SIZE), dtype=np.int32)
    num_splits = np.zeros(batch_size + 1, dtype=np.int32)
    now_idx = now_block = now_n_split_idx = cum_num_splits = 0
    for i in range(num_sm_parts):
        meta[i, 0] = now_idx
        meta[i, 1] = now_block * block_size_n
        meta[i, 4] = now_n_split_idx
        remain_payload = payload
        while now_idx < batch_size:
            nb = num_blocks[now_idx]
            now_remain_blocks = nb - now_block
            if remain_payload >= now_remain_blocks + fixed:
                cum_num_splits += now_n_split_idx + 1
                num_splits[now_idx + 1] = cum_num_splits
                remain_payload -= now_remain_blocks + fixed
                now_idx += 1
                now_block = 0
                now_n_split_idx = 0
            else:
                if remain_payload - fixed > 0:
                    now_block += remain_payload - fixed
                    now_n_split_idx += 1
                    remain_payload = 0
                break
        meta[i, 2] = now_idx if now_block > 0 else now_idx - 1
        meta[i, 3] = now_block * block_size_n if now_block > 0 else int(seqlens_k[now_idx - 1])
    assert now_idx == batch_size and now_block == 0 and now_n_split_idx == 0
    return meta, num_splits


# --------------------------------------------------------------------------- #
# a-6 / a-7 / a-8  quantisers                                                   #
# --------------------------------------------------------------------------- #
def quantize_rows(x_bf16_as_f32: np.ndarray, head_dim_v: int = HEAD_DIM_V):
    """One warp per row: csrc/kernels/ckv_fp8_per_token/cache_kernels.cu:27-98.

    x: float32 array [..., D] whose values are exactly representable in bf16.
    amax over the first head_dim_v lanes is taken in bf16 (:70-75; exact, max of
    bf16 values), scale = amax/448 (:76), scale = fma(scale, 1, 1e-12) (:78),
    inv = 1/scale (:79), nope -> e4m3(float(x)*inv) (:84), rope ->
    bf16_rn(float(x)*inv) (:94).  Returns (codes u8 [...,dv], scale f32 [...],
    rope f32 [..., D-dv] holding bf16 values).
    """
    x = np.asarray(x_bf16_as_f32, dtype=np.float32)
    nope = x[..., :head_dim_v]
    rope = x[..., head_dim_v:]
    amax = np.max(np.abs(nope), axis=-1).astype(np.float32)
    scale = (amax / FP8_MAX).astype(np.float32)
    scale = (scale + np.float32(1e-12)).astype(np.float32)
    inv = (np.float32(1.0) / scale).astype(np.float32)
    codes = f32_to_e4m3((nope * inv[..., None]).astype(np.float32))
    rope_q = round_to_bf16((rope * inv[..., None]).astype(np.float32))
    return codes, scale, rope_q


def quantize_ckv_per_token_head(q: np.ndarray, head_dim_v: int = HEAD_DIM_V):
    """csrc/flash_api.cpp:399-460 -> (nope codes [...,dv], scale [...,1], rope [...,D-dv])."""
    codes, scale, rope = quantize_rows(q, head_dim_v)
    return codes, scale[..., None], rope


def quantize_and_cache_k(key, k_lora_cache, k_scale_cache, k_rope_cache, indices,
                         head_dim_v: int = HEAD_DIM_V):
    """csrc/flash_api.cpp:462-532, cache_kernels.cu:140-218: quantise key rows and
    scatter them to ``indices`` (token slots) in the three pool arrays, in place.
    key [n, 1, D] (bf16 values as f32); caches [S,1,dv] u8 / [S,1,1] f32 / [S,1,D-dv] f32."""
    key = np.asarray(key, dtype=np.float32).reshape(-1, key.shape[-1])
    codes, scale, rope = quantize_rows(key, head_dim_v)
    idx = np.asarray(indices, dtype=np.int64)
    for i, loc in enumerate(idx):  # later duplicates win, as with racing warps' last writer
        k_lora_cache[loc, 0, :] = codes[i]
        k_scale_cache[loc, 0, 0] = scale[i]
        k_rope_cache[loc, 0, :] = rope[i]
    return k_lora_cache, k_scale_cache, k_rope_cache


def dequantize_ckv_fused_indexed(k_lora_cache, k_rope_cache, k_scale_cache, indices):
    """cache_kernels.cu:274-369: gather ``indices`` and dequantise:
    lora = bf16_rn(float(e4m3)*scale), rope = bf16_rn(float(rope)*scale)."""
    idx = np.asarray(indices, dtype=np.int64)
    scale = np.asarray(k_scale_cache, dtype=np.float32).reshape(-1)[idx]
    lora = e4m3_to_f32(np.asarray(k_lora_cache)[idx, 0, :])
    rope = np.asarray(k_rope_cache, dtype=np.float32)[idx, 0, :]
    lora_deq = round_to_bf16((lora * scale[:, None]).astype(np.float32))
    rope_deq = round_to_bf16((rope * scale[:, None]).astype(np.float32))
    return lora_deq[:, None, :], rope_deq[:, None, :]


# --------------------------------------------------------------------------- #
# a-1  split-KV attention, faithful flavour                                     #
# --------------------------------------------------------------------------- #
def _mask_len(s_q: int, q_head_per_hk: int, q_seq_per_hk: int, global_row: int) -> int:
    """splitkv_mla.cu:1136-1145 (get_mask_len)."""
    if global_row < q_seq_per_hk:
        return s_q - global_row // q_head_per_hk - 1
    return 0


def _exp2(x):
    return np.exp2(x.astype(np.float32)).astype(np.float32)


def _partial_attention(qn, qr, qs, kn_pages, kr_pages, ks_pages, block_ids, first_token,
                       seqlen_k, right_border, scale_log2):
    """One (request, m-block, sm-part) segment: sequential page-by-page online softmax
    with per-row per-page FP8 re-quantisation of P.

    Follows wg0_bunch_0 / wg1_bunch_0 (splitkv_mla.cu:377-458, 473-574); the two
    warpgroups' seesaw is arithmetically the plain sequence below (each page: one
    (m, p_scale) update per row; O and L rescaled by
    exp2(m_old-m_new)*p_old/p_new, :442-455).

    qn [M,512] f32 (decoded e4m3), qr [M,64] f32 (bf16 values), qs [M] f32.
    kn_pages u8 [P,64,512], kr_pages f32 [P,64,64], ks_pages f32 [P,64].
    block_ids: page ids for the blocks of this segment, in order.
    first_token: token index of the first block's first row.
    Returns (O [M,512] un-normalised in units of p, L [M], m [M], p [M]).
    """
    M = qn.shape[0]
    m = np.full(M, MAX_INIT_VAL_SM, dtype=np.float32)       # :1300
    p = np.ones(M, dtype=np.float32)                         # :1301
    L = np.zeros(M, dtype=np.float32)
    O = np.zeros((M, HEAD_DIM_V), dtype=np.float32)
    tok0 = first_token
    for pid in block_ids:
        kn = e4m3_to_f32(kn_pages[pid])                      # [64,512]
        kr = kr_pages[pid]
        ks = ks_pages[pid].astype(np.float32)
        tok = tok0 + np.arange(PAGE_BLOCK_SIZE)
        valid_row = tok < seqlen_k
        # fill_oob_V (:698-721): rows at/after seqlen_k are zeroed before use as V;
        # as K they only feed masked columns.
        kn_v = np.where(valid_row[:, None], kn, np.float32(0.0))
        with np.errstate(invalid="ignore", over="ignore"):
            s = (qn @ np.where(valid_row[:, None], kn, 0.0).T + qr @ np.where(valid_row[:, None], kr, 0.0).T).astype(np.float32)
        ks_ok = np.where(valid_row, ks, np.float32(0.0))
        x = (s * (qs[:, None] * ks_ok[None, :]).astype(np.float32)).astype(np.float32)   # :400-401
        in_border = tok[None, :] < right_border[:, None]                                    # :403-407
        x = np.where(in_border, x, MAX_INIT_VAL)
        ks_eff = np.where(in_border, ks_ok[None, :], np.float32(0.0)).astype(np.float32)
        cur_max = np.maximum(np.max(x, axis=1), MAX_INIT_VAL).astype(np.float32)            # :396-412
        cur_max = (cur_max * scale_log2).astype(np.float32)                                 # :415
        new_max = np.maximum(m, cur_max)                                                    # :416
        e = _exp2((x * scale_log2).astype(np.float32) - new_max[:, None])                   # :424-425
        cur_sum = np.sum(e, axis=1, dtype=np.float32)                                       # :426
        pp = (e * ks_eff).astype(np.float32)                                                # :427-428
        cur_pmax = np.maximum(np.max(pp, axis=1), MAX_INIT_VAL_P).astype(np.float32)        # :420,429-432
        p_scale = (cur_pmax / FP8_MAX).astype(np.float32)                                   # :433
        cur_sum = (cur_sum / p_scale).astype(np.float32)                                    # :434
        p8 = e4m3_to_f32(f32_to_e4m3((pp / p_scale[:, None]).astype(np.float32)))           # :436-440
        scale_for_old = (_exp2(m - new_max) * p / p_scale).astype(np.float32)               # :442
        O = (O * scale_for_old[:, None] + p8 @ kn_v).astype(np.float32)                     # :449-454 + PV gemm
        L = (L * scale_for_old + cur_sum).astype(np.float32)                                # :455
        m, p = new_max, p_scale                                                             # :446-447
        tok0 += PAGE_BLOCK_SIZE
    return O, L, m, p


def combine_partials(o_accum, lse_accum, num_splits, out, lse, only_requests=None):
    """a-5, ckv_fp8_per_token/mla_combine.cu:16-154: merge the split-KV partials of every request that has more than one
    split.  o_accum f32 [splits, M, dv]; lse_accum f32 [splits, M] (log2 domain); num_splits i32 [b + 1]; writes
    out f32 [b, M, dv] (bf16 values) and lse f32 [b, M] (natural log) in place; single-split requests are untouched (:31-33).
    Pinned against the reference kernel's source run on CPU (tests/golden/ref_kernels_cpu.npz)."""
    b, M, dv = out.shape
    for batch_idx in range(b):
        if only_requests is not None and batch_idx not in only_requests:
            continue
        s0, s1 = int(num_splits[batch_idx]), int(num_splits[batch_idx + 1])
        if s1 - s0 == 1:
            continue                                             # :31-33
        lse_i = lse_accum[s0:s1]                                 # [n, M] (log2 domain)
        max_lse = np.max(lse_i, axis=0)
        max_lse = np.where(np.isneginf(max_lse), np.float32(0.0), max_lse)       # :92
        with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
            sum_lse = np.sum(_exp2(lse_i - max_lse[None, :]), axis=0, dtype=np.float32)
            bad = (sum_lse == 0.0) | np.isnan(sum_lse)
            global_lse = np.where(bad, np.float32(np.inf), np.log2(sum_lse).astype(np.float32) + max_lse)  # :101
            lse[batch_idx] = (global_lse / LOG2E).astype(np.float32)              # :103
            w = _exp2(lse_i - global_lse[None, :])                                # :108
        acc = np.zeros((M, dv), dtype=np.float32)
        for i in range(s1 - s0):                                                  # :128-137
            wi = w[i]
            with np.errstate(invalid="ignore"):
                contrib = np.where((wi != 0.0)[:, None], wi[:, None] * o_accum[s0 + i], np.float32(0.0))  # :131 skip w==0
            acc = (acc + contrib).astype(np.float32)
        out[batch_idx] = round_to_bf16(acc)


def flash_mla_ckv_fp8_per_token(q_nope, q_rope, k_cache_lora, k_cache_rope, q_scale, k_scale,
                                block_table, cache_seqlens, tile_scheduler_metadata, num_splits,
                                softmax_scale=None, causal=False, block_m: int = BLOCK_SIZE_M,
                                return_partials: bool = False, only_requests=None):
    """Faithful restatement of ``flash_mla_ckv_fp8_per_token``
    (flash_mla_fp8/flash_mla_interface.py:68-120 -> csrc/flash_api.cpp:216-397 ->
    splitkv_mla.cu:1147-1515 + mla_combine.cu:16-154).

    q_nope u8 [b,s_q,h,512] (e4m3 codes); q_rope f32 [b,s_q,h,64] (bf16 values);
    q_scale f32 [b,s_q,h,1]; k_cache_lora u8 [pages,64,1,512]; k_cache_rope f32
    [pages,64,1,64]; k_scale f32 [pages,64,1,1]; block_table i32 [b,max_pages];
    cache_seqlens i32 [b]; metadata from get_mla_metadata.
    Returns (out f32 [b,s_q,h,512] holding bf16 values, lse f32 [b,h,s_q]).
    only_requests: optional set of batch indices to evaluate (full-size tests sample a few requests; the others' rows of
    the result are meaningless).
    """
    q_nope = np.asarray(q_nope)
    b, s_q, h_q, dv = q_nope.shape
    assert dv == HEAD_DIM_V
    h_k = k_cache_lora.shape[2]
    assert h_k == 1, "MLA has one latent KV head"
    if softmax_scale is None:
        softmax_scale = (dv + q_rope.shape[-1]) ** (-0.5)      # flash_mla_interface.py:100-101
    if s_q == 1:
        causal = False                                          # flash_api.cpp:279
    scale_log2 = np.float32(np.float32(softmax_scale) * np.float32(math.log2(math.e)))  # :324
    M = s_q * h_q                                               # q_seq_per_hk, flash_api.cpp:282
    qn = e4m3_to_f32(q_nope.reshape(b, M, dv))
    qr = np.asarray(q_rope, dtype=np.float32).reshape(b, M, -1)
    qs = np.asarray(q_scale, dtype=np.float32).reshape(b, M)
    kn_pages = np.asarray(k_cache_lora)[:, :, 0, :]
    kr_pages = np.asarray(k_cache_rope, dtype=np.float32)[:, :, 0, :]
    ks_pages = np.asarray(k_scale, dtype=np.float32)[:, :, 0, 0]
    meta = np.asarray(tile_scheduler_metadata)
    num_splits = np.asarray(num_splits)
    num_sm_parts = meta.shape[0]
    total_num_splits = b + num_sm_parts                         # flash_api.cpp:365
    num_m_blocks = ceil_div(M, block_m)

    out = np.zeros((b, M, dv), dtype=np.float32)
    lse = np.zeros((b, M), dtype=np.float32)
    o_accum = np.full((total_num_splits, M, dv), np.nan, dtype=np.float32)
    lse_accum = np.full((total_num_splits, M), np.nan, dtype=np.float32)

    for part in range(num_sm_parts):
        begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split_idx = (int(v) for v in meta[part, :5])
        if begin_idx >= b:
            continue                                            # splitkv_mla.cu:1226
        for batch_idx in range(begin_idx, end_idx + 1):
            if only_requests is not None and batch_idx not in only_requests:
                continue
            n_split_idx = begin_n_split_idx if batch_idx == begin_idx else 0
            seqlen_k = int(cache_seqlens[batch_idx])
            start_block_idx = begin_seqlen // PAGE_BLOCK_SIZE if batch_idx == begin_idx else 0
            end_block_idx0 = (ceil_div(end_seqlen, PAGE_BLOCK_SIZE) if batch_idx == end_idx
                              else ceil_div(seqlen_k, PAGE_BLOCK_SIZE))
            is_no_split = start_block_idx == 0 and end_block_idx0 == ceil_div(seqlen_k, PAGE_BLOCK_SIZE)  # :1239
            for mb in range(num_m_blocks):
                r0, r1 = mb * block_m, min((mb + 1) * block_m, M)
                rows = np.arange(r0, r1)
                end_block_idx = end_block_idx0
                if causal:                                       # :1242-1266
                    common_mask_len = _mask_len(s_q, h_q // h_k, M, mb * block_m + block_m - 1)
                    lim = seqlen_k - common_mask_len
                    if batch_idx == end_idx:
                        lim = min(end_seqlen, lim)
                    end_block_idx = ceil_div(max(lim, 0), PAGE_BLOCK_SIZE)
                    border = np.array([min(seqlen_k - _mask_len(s_q, h_q // h_k, M, int(r)),
                                           end_block_idx * PAGE_BLOCK_SIZE) for r in rows], dtype=np.int64)
                else:
                    border = np.full(rows.shape, seqlen_k, dtype=np.int64)
                blocks = [int(block_table[batch_idx, j]) for j in range(start_block_idx, end_block_idx)]
                O, L, m, p = _partial_attention(
                    qn[batch_idx, r0:r1], qr[batch_idx, r0:r1], qs[batch_idx, r0:r1],
                    kn_pages, kr_pages, ks_pages, blocks, start_block_idx * PAGE_BLOCK_SIZE,
                    seqlen_k, border, scale_log2)
                bad = (L == 0.0) | np.isnan(L)                   # :1464-1466
                L_safe = np.where(bad, np.float32(1.0), L)
                o_norm = (O / L_safe[:, None]).astype(np.float32)  # :756 / :795
                with np.errstate(divide="ignore", invalid="ignore"):
                    if is_no_split:
                        out[batch_idx, r0:r1] = round_to_bf16(o_norm)
                        lse[batch_idx, r0:r1] = np.where(
                            bad, np.float32(np.inf),
                            (np.log((L * p).astype(np.float32)) + m / LOG2E).astype(np.float32))  # :1484
                    else:
                        split_idx = int(num_splits[batch_idx]) + n_split_idx                      # :1490
                        o_accum[split_idx, r0:r1] = o_norm
                        lse_accum[split_idx, r0:r1] = np.where(
                            bad, np.float32(-np.inf),
                            (np.log2((L * p).astype(np.float32)) + m).astype(np.float32))          # :1506

    # ---- combine (mla_combine.cu:16-154) ----
    combine_partials(o_accum, lse_accum, num_splits, out, lse, only_requests)

    out = out.reshape(b, s_q, h_q, dv)
    lse = lse.reshape(b, s_q, h_q).transpose(0, 2, 1).copy()     # flash_api.cpp:392-394
    if return_partials:
        return out, lse, o_accum, lse_accum
    return out, lse


# --------------------------------------------------------------------------- #
# a-1 semantic flavour and a-2 (bf16) : plain fp32/fp64 attention               #
# --------------------------------------------------------------------------- #
def _dense_attention(q, k, v, softmax_scale, s_q, h_q, causal):
    """q [M,d] f64 ordered (s_q, h); k [S,d]; v [S,dv].  Mirrors the test reference
    tests/test_flash_mla_ckv_fp8_per_token.py:14-30 (tril(diagonal=s_k-s_q))."""
    S = k.shape[0]
    logits = (q @ k.T) * softmax_scale
    if causal and s_q > 1:
        sq_idx = np.arange(q.shape[0]) // h_q
        allowed = np.arange(S)[None, :] <= (S - s_q + sq_idx)[:, None]
        logits = np.where(allowed, logits, -np.inf)
    if S == 0:
        return np.zeros((q.shape[0], v.shape[1])), np.full(q.shape[0], np.inf)
    mx = np.max(logits, axis=1, keepdims=True)
    mx = np.where(np.isfinite(mx), mx, 0.0)
    ex = np.exp(logits - mx)
    den = np.sum(ex, axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        o = np.where(den > 0, (ex @ v) / np.where(den > 0, den, 1.0), 0.0)
        lse = np.where(den[:, 0] > 0, np.log(np.where(den[:, 0] > 0, den[:, 0], 1.0)) + mx[:, 0], np.inf)
    return o, lse


def flash_mla_ckv_fp8_semantic(q_nope, q_rope, k_cache_lora, k_cache_rope, q_scale, k_scale,
                               block_table, cache_seqlens, softmax_scale=None, causal=False):
    """Semantic flavour: dequantise everything, float64 softmax attention.  What the
    kernel computes if P were not re-quantised."""
    q_nope = np.asarray(q_nope)
    b, s_q, h_q, dv = q_nope.shape
    if softmax_scale is None:
        softmax_scale = (dv + q_rope.shape[-1]) ** (-0.5)
    M = s_q * h_q
    qs = np.asarray(q_scale, dtype=np.float64).reshape(b, M, 1)
    qn = e4m3_to_f32(q_nope.reshape(b, M, dv)).astype(np.float64) * qs
    qr = np.asarray(q_rope, dtype=np.float64).reshape(b, M, -1) * qs
    out = np.zeros((b, M, dv))
    lse = np.zeros((b, M))
    for i in range(b):
        n = int(cache_seqlens[i])
        nblk = ceil_div(n, PAGE_BLOCK_SIZE)
        pages = np.asarray(block_table[i, :nblk], dtype=np.int64)
        ks = np.asarray(k_scale, dtype=np.float64)[pages, :, 0, 0].reshape(-1)[:n]
        kn = e4m3_to_f32(np.asarray(k_cache_lora)[pages, :, 0, :]).reshape(-1, dv)[:n].astype(np.float64) * ks[:, None]
        kr = np.asarray(k_cache_rope, dtype=np.float64)[pages, :, 0, :].reshape(-1, q_rope.shape[-1])[:n] * ks[:, None]
        o, l = _dense_attention(np.concatenate([qn[i], qr[i]], axis=1), np.concatenate([kn, kr], axis=1),
                                kn, softmax_scale, s_q, h_q, causal)
        out[i], lse[i] = o, l
    return (out.reshape(b, s_q, h_q, dv).astype(np.float32),
            lse.reshape(b, s_q, h_q).transpose(0, 2, 1).astype(np.float32))


def flash_mla_with_kvcache(q, k_cache, block_table, cache_seqlens, head_dim_v=HEAD_DIM_V,
                           softmax_scale=None, causal=False):
    """a-2: bf16 Q / bf16 paged K (flash_mla_interface.py:26-66 -> flash_api.cpp:65-214).
    float64 softmax attention over the gathered pages; out rounded to bf16."""
    q = np.asarray(q, dtype=np.float64)
    b, s_q, h_q, d = q.shape
    if softmax_scale is None:
        softmax_scale = d ** (-0.5)
    M = s_q * h_q
    out = np.zeros((b, M, head_dim_v))
    lse = np.zeros((b, M))
    for i in range(b):
        n = int(cache_seqlens[i])
        nblk = ceil_div(n, PAGE_BLOCK_SIZE)
        pages = np.asarray(block_table[i, :nblk], dtype=np.int64)
        k = np.asarray(k_cache, dtype=np.float64)[pages, :, 0, :].reshape(-1, d)[:n]
        o, l = _dense_attention(q[i].reshape(M, d), k, k[:, :head_dim_v], softmax_scale, s_q, h_q, causal)
        out[i], lse[i] = o, l
    return (round_to_bf16(out.reshape(b, s_q, h_q, head_dim_v).astype(np.float32)),
            lse.reshape(b, s_q, h_q).transpose(0, 2, 1).astype(np.float32))
