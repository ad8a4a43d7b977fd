"""ctypes binding of the in-tree C-ABI library (include/b200_decode.h).

There is no CPU or eager fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_DECODE_LIB") or os.path.join(_HERE, "_C", "libb200decode.so")  # override: A/B of two builds

_P = c_void_p
_SIGNATURES = {
    "b200_last_error": (c_char_p, []),
    "b200_sm_count": (c_int, []),
    "b200_mla_num_sm_parts": (c_int, [c_int, c_int, c_int]),
    "b200_mla_get_metadata": (c_int, [_P, c_int, c_int, _P, _P, _P]),
    "b200_mla_get_metadata_ex": (c_int, [_P, c_int, c_int, c_int, _P, _P, _P]),
    "b200_mla_quantize_q": (c_int, [_P, c_int64, c_int, c_int, _P, _P, _P, _P]),
    "b200_mla_quantize_and_cache_k": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, _P]),
    "b200_mla_rope_quantize_fused": (
        c_int,
        [_P, c_int64, c_int64, _P, c_int64, c_int64, _P, c_int64, _P, _P, c_int, c_int, c_int64, c_int, _P, _P, _P, _P,
         _P, _P, _P, _P],
    ),
    "b200_mla_dequantize_indexed": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, _P, _P]),
    "b200_mla_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "b200_mla_merge_state_ints": (c_size_t, [c_int, c_int, c_int]),
    "b200_mla_decode_ckv_fp8_per_token": (
        c_int,
        [_P, _P, _P, _P, _P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P,
         _P, c_size_t, _P, c_int, _P],
    ),
    "b200_mla_decode_bf16": (
        c_int,
        [_P, _P, c_int64, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, c_int, c_float, c_int, _P, _P, _P, c_size_t,
         c_int, _P],
    ),
    "b200_mla_combine": (c_int, [_P, _P, _P, _P, _P, c_int, c_int, c_int, c_int, _P]),
    "b200_comm_flag_bytes": (c_size_t, []),
    "b200_comm_alloc": (c_int, [c_size_t, _P]),
    "b200_comm_free": (c_int, [_P]),
    "b200_comm_ipc_export": (c_int, [_P, _P]),
    "b200_comm_ipc_open": (c_int, [_P, _P]),
    "b200_comm_ipc_close": (c_int, [_P]),
    "b200_comm_all_reduce_bf16": (c_int, [_P, _P, c_size_t, c_int, c_int, _P, c_int64, _P, c_int, _P]),
    "b200_comm_all_reduce_nvls_bf16": (c_int, [_P, _P, _P, c_size_t, c_int, c_int, _P, c_int64, _P, c_int, _P]),
    "b200_comm_reduce_scatter_bf16": (c_int, [_P, _P, c_size_t, c_int, c_int, _P, _P, c_int64, _P, c_int, _P]),
    "b200_comm_all_gather_bf16": (c_int, [_P, _P, c_size_t, c_int, c_int, _P, _P, c_int64, _P, c_int, _P]),
    "b200_act_quant_1x128": (c_int, [_P, c_int64, c_int64, _P, _P, c_int64, c_float, _P]),
    "b200_act_quant_1x128_strided": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, _P, _P, c_int64, c_float, _P]),
    "b200_qk_rmsnorm_quant": (c_int, [_P, c_int64, c_int64, c_int, c_int, _P, _P, c_float, _P, c_int64, _P, _P, c_int64, c_float, _P]),
    "b200_fp8_gemm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "b200_fp8_gemm_nt": (c_int, [_P, _P, c_int64, _P, _P, _P, c_int, c_int, c_int, _P, c_size_t, c_int, _P]),
    "b200_fp8_gemm_grouped_row_tiles": (c_int, [c_int, c_int, c_int]),
    "b200_fp8_gemm_grouped_nt": (c_int, [_P, _P, c_int64, _P, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int, _P]),
    "b200_mla_read_trace": (c_int, [_P, c_int]),
    "b200_mla_read_part_times": (c_int, [_P, c_int]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)

# test-only hardware probes (include/b200_lab.h, _C/libb200lab.so): never loaded by the product modules
LAB_LIB_PATH = os.path.join(_HERE, "_C", "libb200lab.so")
_LAB_SIGNATURES = {
    "b200_lab_last_error": (c_char_p, []),
    "b200_umma_selftest": (c_int, [c_int, c_int, ctypes.POINTER(c_float)]),
    "b200_umma_rate": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_float)]),
    "b200_umma_pair_rate": (c_int, [c_int, c_int, c_int, ctypes.POINTER(c_float)]),
    "b200_umma_pair_lab": (c_int, [c_int, c_int, _P, _P, _P]),
}
LAB_EXPORTED_SYMBOLS = tuple(_LAB_SIGNATURES)
_lab = None


def lab() -> ctypes.CDLL:
    """The probe library (tests / tools only)."""
    global _lab
    if _lab is None:
        if not os.path.exists(LAB_LIB_PATH):
            raise RuntimeError("sglang_fluentllm_b200: %s not found -- run `python __graft_entry__.py`" % LAB_LIB_PATH)
        handle = ctypes.CDLL(LAB_LIB_PATH)
        for name, (res, args) in _LAB_SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lab = handle
    return _lab

_lib = None


def lib() -> ctypes.CDLL:
    """Load (once) and return the library; raise loudly when it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "sglang_fluentllm_b200: %s not found -- run `python __graft_entry__.py` (nvcc, sm_100a) first; "
                "there is no CPU fallback for this path" % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().b200_last_error()
        raise RuntimeError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))


def ptr(t) -> int:
    return t.data_ptr()


def current_stream(device=None) -> int:
    import torch

    return torch.cuda.current_stream(device).cuda_stream
