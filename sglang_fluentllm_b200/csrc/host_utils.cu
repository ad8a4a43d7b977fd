#include "host_utils.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

namespace b200 {

static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}
const char* get_last_error() { return g_last_error; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is fetched through the runtime so the library has no link-time libcuda dependency
// (the reference's DeepGEMM does the same: deep_gemm/include/deep_gemm/tma_utils.cuh:43-58).
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tma_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t cols, uint64_t rows, uint64_t pitch_bytes,
                uint32_t box_cols, uint32_t box_rows, TmaSwizzle swz) {
  EncodeTiledFn fn = get_encode_fn();
  B200_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  CUtensorMapDataType dt;
  switch (elem_bytes) {
    case 1: dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; break;
    case 2: dt = CU_TENSOR_MAP_DATA_TYPE_UINT16; break;
    case 4: dt = CU_TENSOR_MAP_DATA_TYPE_UINT32; break;
    default: set_last_error("make_tma_2d: unsupported element size %d", elem_bytes); return 1;
  }
  CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_NONE;
  if (swz == TmaSwizzle::B32) sw = CU_TENSOR_MAP_SWIZZLE_32B;
  if (swz == TmaSwizzle::B64) sw = CU_TENSOR_MAP_SWIZZLE_64B;
  if (swz == TmaSwizzle::B128) sw = CU_TENSOR_MAP_SWIZZLE_128B;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dt, 2, const_cast<void*>(gptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B200_CHECK(r == CUDA_SUCCESS,
             "cuTensorMapEncodeTiled failed (%d): ptr=%p cols=%llu rows=%llu pitch=%llu box=%ux%u elem=%d", (int)r, gptr,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_bytes, box_cols, box_rows,
             elem_bytes);
  return 0;
}

bool pdl_default() {
  static const bool on = getenv("B200_NO_PDL") == nullptr;
  return on;
}

int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace b200
