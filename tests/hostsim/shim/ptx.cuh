// TEST INFRASTRUCTURE ONLY: host stand-ins for the handful of csrc/ptx.cuh wrappers the streaming kernels of
// csrc/small_kernels.cu use, so that file's kernels can be compiled by g++ and run on CPU threads
// (tests/hostsim/build_hostsim.py).  Found before the real ptx.cuh on the include path of that build only.
#pragma once
#include <cmath>
#include <cstdint>
namespace b200 {
inline void pdl_wait_prior_grid() {}     // griddepcontrol.wait: kernels run one after the other here
inline void pdl_launch_dependents() {}   // griddepcontrol.launch_dependents
// cvt.rn.satfinite.e4m3x2.f32 (lo = a, hi = b) through NVIDIA's host conversion
inline uint32_t cvt_e4m3x2(float a, float b) {
  return (uint32_t)__nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3) |
         ((uint32_t)__nv_cvt_float_to_fp8(b, __NV_SATFINITE, __NV_E4M3) << 8);
}
inline uint32_t cvt_e4m3x4(float a, float b, float c, float d) { return cvt_e4m3x2(a, b) | (cvt_e4m3x2(c, d) << 16); }
// redux.sync.max.f32 over the full warp (NaN operands are ignored, as fmaxf does)
inline float warp_max_f32(float v) {
  for (int o = 16; o > 0; o >>= 1) v = std::fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
}  // namespace b200

// rsqrt.approx stand-in for the host build (qk_rmsnorm_quant_kernel)
inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
