// TEST INFRASTRUCTURE ONLY (oracle/): a host-side stand-in for the CUDA execution model, so that the reference's own
// kernel sources (3rdparty/flashmla-fp8/csrc/kernels/get_mla_metadata.cu, ckv_fp8_per_token/cache_kernels.cu) compile
// with g++ and run on CPU threads.  Pre-included (-include) in front of the reference sources by build_ref.py.
// One std::thread per CUDA thread of a block; blocks run one after the other; warp shuffles and __syncthreads are
// barriers + an exchange buffer.  The arithmetic is the reference's, through NVIDIA's own host implementations of
// __nv_bfloat16 / __nv_fp8_e4m3 (cuda_bf16.h, cuda_fp8.h).
#pragma once
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <cuda_runtime_api.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#undef __global__
#undef __device__
#undef __host__
#undef __forceinline__
#undef __launch_bounds__
#undef __grid_constant__
#undef __shared__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static /* a block's shared variable: blocks run one at a time (extern arrays: see build_ref.py) */

namespace refshim {
struct Idx { unsigned x = 0, y = 0, z = 0; };
struct Block {
  explicit Block(int n) : threads(n), block_bar(n) {
    for (int w = 0; w < (n + 31) / 32; ++w) warp_bar.emplace_back(std::make_unique<std::barrier<>>(std::min(32, n - w * 32)));
    xch.resize((size_t)n);
  }
  int threads;
  std::barrier<> block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> xch;
};
extern thread_local Block* cur;
void run_grid(unsigned grid_x, unsigned block_x, const std::function<void()>& kernel, unsigned grid_y = 1);
}  // namespace refshim

extern thread_local refshim::Idx threadIdx, blockIdx, blockDim, gridDim;

template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane, int = 32) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  refshim::Block& b = *refshim::cur;
  const unsigned tid = threadIdx.x, warp = tid / 32;
  uint64_t raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  b.xch[tid] = raw;
  b.warp_bar[warp]->arrive_and_wait();
  const uint64_t got = b.xch[warp * 32 + (unsigned)src_lane];
  b.warp_bar[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int w = 32) {
  return __shfl_sync(m, v, (int)((threadIdx.x % 32) ^ (unsigned)lane_mask), w);
}
inline void __syncwarp(unsigned = 0xffffffffu) { refshim::cur->warp_bar[threadIdx.x / 32]->arrive_and_wait(); }
inline void __syncthreads() { refshim::cur->block_bar.arrive_and_wait(); }
template <class T>
inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
// IEEE single-precision operations with round-to-nearest (the host build is compiled with -ffp-contract=off)
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
// bit casts and the two atomics the (uninstantiated-here) per-tensor kernels of sgl_fp8_quantization.cu name
inline float __int_as_float(int v) { float f; std::memcpy(&f, &v, 4); return f; }
inline int __float_as_int(float f) { int v; std::memcpy(&v, &f, 4); return v; }
inline float __uint_as_float(unsigned v) { float f; std::memcpy(&f, &v, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned v; std::memcpy(&v, &f, 4); return v; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; *p = std::max(o, v); return o; }  // blocks run one at a time;
template <class T> inline T atomicMin(T* p, T v) { T o = *p; *p = std::min(o, v); return o; }  // not used by the built kernels
// programmatic dependent launch has no meaning when kernels run one after the other
#define cudaGridDependencySynchronize() ((void)0)
#define cudaTriggerProgrammaticLaunchCompletion() ((void)0)
using std::max;
using std::min;
