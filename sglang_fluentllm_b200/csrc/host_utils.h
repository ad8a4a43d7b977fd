// Host-side helpers shared by the C-ABI entry points: error reporting and TMA tensor-map construction.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace b200 {

// Thread-local last-error string returned by b200_last_error().
void set_last_error(const char* fmt, ...);
const char* get_last_error();

#define B200_CHECK(cond, ...)            \
  do {                                   \
    if (!(cond)) {                       \
      ::b200::set_last_error(__VA_ARGS__); \
      return 1;                          \
    }                                    \
  } while (0)

#define B200_CUDA(expr)                                                                        \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      ::b200::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                                \
    }                                                                                          \
  } while (0)

enum class TmaSwizzle { None, B32, B64, B128 };

// 2-D row-major tensor: `rows` x `row_bytes_or_elems` elements of `elem_bytes`, row pitch `pitch_bytes`.
// Box = box_cols elements x box_rows rows.  Returns 0 on success.
int make_tma_2d(CUtensorMap* out, const void* gptr, int elem_bytes, uint64_t cols, uint64_t rows, uint64_t pitch_bytes,
                uint32_t box_cols, uint32_t box_rows, TmaSwizzle swz);

int device_sm_count();

// Programmatic dependent launch for the path's small kernels (quantisers): on unless B200_NO_PDL is set.
bool pdl_default();

// Launch with (optionally) the programmatic-stream-serialization attribute: the kernel may be scheduled while its
// predecessor in the stream drains; it must execute griddepcontrol.wait before touching the predecessor's outputs.
template <class... KArgs, class... Args>
int launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                  Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  return 0;
}

}  // namespace b200
