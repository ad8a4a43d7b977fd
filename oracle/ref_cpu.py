"""The reference's own quantiser / metadata kernels executed on CPU (TEST INFRASTRUCTURE ONLY).

``oracle/_ref/libflashmla_ref_cpu.so`` is built by ``oracle/ref_build/build_ref.py`` from the kernel sources under
``/root/reference/3rdparty/flashmla-fp8/csrc/kernels`` (get_mla_metadata.cu, ckv_fp8_per_token/cache_kernels.cu),
compiled unmodified for the host against a thread/warp shim.  It pins the restatements in oracle/mla.py bit for bit
(tests/test_oracle_golden.py, fixture tests/golden/ref_kernels_cpu.npz for machines without the library)."""
import ctypes
import os

import numpy as np

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libflashmla_ref_cpu.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(LIB_PATH)
        P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
        for name, args in {
            "ref_get_mla_metadata": [P, I, I, I, I, P, P],
            "ref_quantize_ckv_per_token_head": [P, I, I, I, P, P, P],
            "ref_quantize_and_cache_k": [P, P, P, P, P, I, I, I, I, I, I],
            "ref_dequantize_fused_indexed": [P, P, P, P, P, P, I, I, I, I, I],
            "ref_per_token_group_quant_fp8": [P, L, L, P, P, L, F],
            "ref_mla_combine": [P, P, P, P, P, I, I, I, I],
        }.items():
            getattr(_lib, name).argtypes = args
            getattr(_lib, name).restype = I
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _aligned(shape, dtype, align=64):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.zeros(n + align, np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n].view(dtype).reshape(shape)


def get_mla_metadata(seqlens, num_heads_per_head_k, num_heads_k, sm_count, block_m=64, block_n=64, fixed_overhead=5):
    """csrc/flash_api.cpp:27-63 + get_mla_metadata.cu:8-80.  Returns (metadata int32 [parts, 8] with columns 5-7 zero --
    the kernel leaves them unwritten --, num_splits int32 [b + 1])."""
    seqlens = np.ascontiguousarray(seqlens, np.int32)
    parts = sm_count // num_heads_k // ((num_heads_per_head_k + block_m - 1) // block_m)
    meta = _aligned((parts, 8), np.int32)
    ns = np.zeros(len(seqlens) + 1, np.int32)
    rc = lib().ref_get_mla_metadata(_p(seqlens), len(seqlens), block_n, fixed_overhead, parts, _p(meta), _p(ns))
    assert rc == 0
    return meta.copy(), ns


def quantize_ckv_per_token_head(x_bf16_bits, head_dim_v=512):
    """cache_kernels.cu:26-98 on uint16 bf16 bit patterns [..., 576] -> (codes u8 [..., 512], scale f32 [..., 1],
    rope bf16 bits u16 [..., 64])."""
    x = np.ascontiguousarray(x_bf16_bits, np.uint16)
    d = x.shape[-1]
    rows = x.size // d
    xa = _aligned(x.shape, np.uint16)
    xa[...] = x
    nope = _aligned(x.shape[:-1] + (head_dim_v,), np.uint8)
    scale = np.zeros(x.shape[:-1] + (1,), np.float32)
    rope = _aligned(x.shape[:-1] + (d - head_dim_v,), np.uint16)
    assert lib().ref_quantize_ckv_per_token_head(_p(xa), rows, d, head_dim_v, _p(nope), _p(scale), _p(rope)) == 0
    return nope.copy(), scale, rope.copy()


def quantize_and_cache_k(key_bf16_bits, k_lora, k_scale, k_rope_bits, loc, head_dim_v=512):
    """cache_kernels.cu:139-218 scattering into caller-owned pool arrays (u8 [slots,1,512], f32 [slots,1,1],
    u16 [slots,1,64]; must be C-contiguous and 16-byte aligned -- use ``pool_arrays``)."""
    key = _aligned(np.shape(key_bf16_bits), np.uint16)
    key[...] = key_bf16_bits
    d = key.shape[-1]
    n = key.size // d
    loc = np.ascontiguousarray(loc, np.int32)
    assert lib().ref_quantize_and_cache_k(_p(key), _p(loc), _p(k_lora), _p(k_scale), _p(k_rope_bits), n, head_dim_v, 1,
                                          d - head_dim_v, d, head_dim_v) == 0


def pool_arrays(slots):
    return _aligned((slots, 1, 512), np.uint8), _aligned((slots, 1, 1), np.float32), _aligned((slots, 1, 64), np.uint16)


def dequantize_ckv_fused_indexed(k_lora, k_rope_bits, k_scale, indices, max_num_tokens=-1):
    """cache_kernels.cu:273-423 -> (lora bf16 bits u16 [n,1,512], rope bf16 bits u16 [n,1,64])."""
    idx = np.ascontiguousarray(indices, np.int32)
    n = len(idx)
    lo, ro = _aligned((n, 1, 512), np.uint16), _aligned((n, 1, 64), np.uint16)
    assert lib().ref_dequantize_fused_indexed(_p(k_lora), _p(k_rope_bits), _p(k_scale), _p(idx), _p(lo), _p(ro), n,
                                              int(max_num_tokens), 512, 1, 64) == 0
    return lo.copy(), ro.copy()


def per_token_group_quant_fp8(x_bf16_bits, eps=1e-10):
    """3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-437 (group 128, column-major TMA-aligned float scales, the
    layout deep_geem.py:86-93 asks for) on uint16 bf16 bits [m, k] -> (codes u8 [m, k], scales f32 [m, k/128])."""
    x = _aligned(np.shape(x_bf16_bits), np.uint16)
    x[...] = x_bf16_bits
    m, k = x.shape
    stride = (m + 3) // 4 * 4
    q = _aligned((m, k), np.uint8)
    sc = np.zeros((k // 128, stride), np.float32)
    assert lib().ref_per_token_group_quant_fp8(_p(x), m, k, _p(q), _p(sc), stride, float(eps)) == 0
    return q.copy(), sc[:, :m].T.copy()


def mla_combine(o_accum, lse_accum, num_splits, batch, s_q, h_q, num_sm_parts):
    """ckv_fp8_per_token/mla_combine.cu:16-154 on fp32 partials [splits, M, 512] / [splits, M] (log2 domain).  Returns
    (out bf16 bits u16 [b, M, 512], lse f32 [b, M]); requests with a single split are left untouched (zeros), :31-33."""
    M = s_q * h_q
    oa = np.ascontiguousarray(o_accum, np.float32)
    la = np.ascontiguousarray(lse_accum, np.float32)
    ns = np.ascontiguousarray(num_splits, np.int32)
    out = _aligned((batch, M, 512), np.uint16)
    lse = np.zeros((batch, M), np.float32)
    assert lib().ref_mla_combine(_p(oa), _p(la), _p(ns), _p(out), _p(lse), batch, s_q, h_q, num_sm_parts) == 0
    return out.copy(), lse
