// TEST INFRASTRUCTURE ONLY: what survives of the TVM-FFI glue once the host-side launcher functions of
// sgl_fp8_quantization.cu are filtered out (one namespace-scope constant built from a DLPack dtype code).
#pragma once
#include <cstdint>
struct RefShimDLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
constexpr RefShimDLDataType dl_int8{0, 8, 1};
constexpr int64_t encode_dlpack_dtype(RefShimDLDataType d) { return (int64_t(d.code) << 16) | (int64_t(d.bits) << 8) | d.lanes; }
