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

}  // namespace b200
