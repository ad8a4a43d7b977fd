"""Rotary embedding + the fused decode prologue (TEST INFRASTRUCTURE ONLY).

Restates, in numpy fp32 with every operation rounded:
  * ``DeepseekScalingRotaryEmbedding.forward_native`` (python/sglang/srt/layers/rotary_embedding.py:804-846):
    ``out = x * cos + rotate(x) * sin`` in fp32 on bf16 inputs, rounded to bf16; ``rotate`` is ``_rotate_gptj`` (:56-60,
    pairs (2i, 2i+1)) or ``_rotate_neox`` (:50-53, pairs (i, i + d/2)); cos/sin rows come from ``cos_sin_cache``
    ``[max_position, rotary_dim]`` = cos | sin (:791-802).  Pinned bit-for-bit by ``tests/golden/rope.npz``, generated from
    that class (tests/golden/make_golden.py).  The CUDA kernel the reference runs on GPU
    (3rdparty/flashinfer/include/flashinfer/pos_enc.cuh:190-208) evaluates the same expression; nvcc may contract one of
    its two products into an FMA, which moves the bf16 result by at most one ulp on a small fraction of elements.
  * the per-layer decode prologue the reference runs as three launches (models/deepseek_v2.py:843-857 ->
    rotary_emb; mem_cache/memory_pool.py:854-871 -> quantize_and_cache_k; layers/attention/flashmla_backend.py:205-206 ->
    quantize_ckv_per_token_head), composed from the restatements in oracle/mla.py.
"""
import numpy as np

from .fp8 import round_to_bf16
from .mla import quantize_and_cache_k, quantize_ckv_per_token_head


def apply_rope(x, positions, cos_sin_cache, is_neox: bool):
    """x: float32 [tokens, heads, rotary_dim] holding bf16 values; positions int [tokens];
    cos_sin_cache float32 [max_position, rotary_dim].  Returns float32 holding the bf16-rounded rotation."""
    x = np.asarray(x, dtype=np.float32)
    d = x.shape[-1]
    cs = np.asarray(cos_sin_cache, dtype=np.float32)[np.asarray(positions, dtype=np.int64)]  # [tokens, d]
    cos, sin = cs[:, : d // 2], cs[:, d // 2:]
    if is_neox:
        cos = np.concatenate([cos, cos], axis=-1)[:, None, :]
        sin = np.concatenate([sin, sin], axis=-1)[:, None, :]
        rot = np.concatenate([-x[..., d // 2:], x[..., : d // 2]], axis=-1)
    else:
        cos = np.repeat(cos, 2, axis=-1)[:, None, :]
        sin = np.repeat(sin, 2, axis=-1)[:, None, :]
        rot = np.stack([-x[..., 1::2], x[..., 0::2]], axis=-1).reshape(x.shape)
    a = (x * cos).astype(np.float32)
    b = (rot * sin).astype(np.float32)
    return round_to_bf16((a + b).astype(np.float32))


def rope_quantize_fused(q_nope, q_pe, kv, positions, cos_sin_cache, is_neox, cache_loc=None, k_lora_cache=None,
                        k_scale_cache=None, k_rope_cache=None):
    """The unfused sequence on the same inputs.  q_nope [t, h, 512], q_pe [t, h, 64], kv [t, 576] (float32 holding
    bf16 values).  Returns (q_nope codes u8 [t, h, 512], q_scale f32 [t, h, 1], q_rope f32 [t, h, 64]) and updates
    the pool arrays ([slots, 1, 512] u8 / [slots, 1, 1] f32 / [slots, 1, 64] f32) in place when cache_loc is given."""
    q_pe_rot = apply_rope(q_pe, positions, cos_sin_cache, is_neox)
    q = np.concatenate([np.asarray(q_nope, dtype=np.float32), q_pe_rot], axis=-1)
    out = quantize_ckv_per_token_head(q)
    if cache_loc is not None:
        kv = np.asarray(kv, dtype=np.float32)
        k_pe_rot = apply_rope(kv[:, None, 512:], positions, cos_sin_cache, is_neox)[:, 0, :]
        key = np.concatenate([kv[:, :512], k_pe_rot], axis=-1)[:, None, :]
        quantize_and_cache_k(key, k_lora_cache, k_scale_cache, k_rope_cache, cache_loc)
    return out
