"""Number-format helpers for the oracle (TEST INFRASTRUCTURE ONLY).

e4m3 ("fn": no inf, max 448, NaN = 0x7f/0xff) with the semantics of the CUDA
conversion the reference kernels use: ``__nv_fp8_e4m3(float)`` /
``cutlass::float_e4m3_t(float)`` = ``cvt.rn.satfinite.e4m3x2.f32`` -- round to
nearest even, saturate finite overflow to +-448
(3rdparty/flashmla-fp8/csrc/kernels/ckv_fp8_per_token/cache_kernels.cu:84,
 splitkv_mla.cu:439).  torch's ``.to(float8_e4m3fn)`` differs only above 464
(NaN instead of saturation); tests check the two agree in range.
"""
import numpy as np

FP8_MAX = np.float32(448.0)


def _e4m3_table():
    vals = np.zeros(256, dtype=np.float64)
    for c in range(256):
        s = -1.0 if c & 0x80 else 1.0
        e = (c >> 3) & 0xF
        m = c & 0x7
        if e == 0xF and m == 0x7:
            vals[c] = np.nan
        elif e == 0:
            vals[c] = s * (m / 8.0) * 2.0 ** (-6)
        else:
            vals[c] = s * (1.0 + m / 8.0) * 2.0 ** (e - 7)
    return vals


E4M3_VALUES_F64 = _e4m3_table()
E4M3_VALUES_F32 = E4M3_VALUES_F64.astype(np.float32)
_POS = E4M3_VALUES_F64[:0x7F]  # 0 .. 448, strictly increasing, codes 0..0x7e


def e4m3_to_f32(codes: np.ndarray) -> np.ndarray:
    """uint8 e4m3 codes -> float32 values (exact)."""
    return E4M3_VALUES_F32[np.asarray(codes, dtype=np.uint8)]


def f32_to_e4m3(x: np.ndarray) -> np.ndarray:
    """float32 -> uint8 e4m3 codes, RNE, satfinite (cvt.rn.satfinite.e4m3x2.f32)."""
    x = np.asarray(x, dtype=np.float32)
    a = np.abs(x).astype(np.float64)
    sign = (np.signbit(x)).astype(np.uint8) << 7
    nan = np.isnan(x)
    a = np.where(nan, 0.0, a)
    sat = a >= 448.0
    a_c = np.minimum(a, 448.0)
    lo = np.searchsorted(_POS, a_c, side="right") - 1  # largest value <= a
    lo = np.clip(lo, 0, 0x7E)
    hi = np.minimum(lo + 1, 0x7E)
    d_lo = a_c - _POS[lo]
    d_hi = _POS[hi] - a_c
    pick_hi = (d_hi < d_lo) | ((d_hi == d_lo) & ((hi & 1) == 0) & (hi != lo))
    code = np.where(pick_hi, hi, lo).astype(np.uint8)
    code = np.where(sat, np.uint8(0x7E), code)
    code = code | sign
    code = np.where(nan, np.uint8(0x7F), code)
    return code.astype(np.uint8)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit pattern (uint16), round-to-nearest-even
    (``__float2bfloat16_rn``, cache_kernels.cu:12-15)."""
    u = np.asarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    nan = np.isnan(np.asarray(x, dtype=np.float32))
    bias = np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))
    r = ((u + bias) >> np.uint64(16)).astype(np.uint16)
    return np.where(nan, np.uint16(0x7FC0), r).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (np.asarray(b, dtype=np.uint16).astype(np.uint32) << np.uint32(16)).view(np.float32)


def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """float32 -> nearest bf16, returned as float32."""
    return bf16_bits_to_f32(f32_to_bf16_bits(x))
