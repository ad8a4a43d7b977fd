// TEST INFRASTRUCTURE ONLY: the slice of flashinfer::vec_t the per-token-group quantiser kernel uses
// (3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:396-434: a 16-byte vector of T filled by cast_load from a T* and
// read element-wise), for the host build of that kernel.
#pragma once
#include <cstddef>
namespace flashinfer {
template <class T, size_t N>
struct vec_t {
  T data[N];
  T& operator[](size_t i) { return data[i]; }
  const T& operator[](size_t i) const { return data[i]; }
  template <class S>
  void cast_load(const S* p) {
    for (size_t i = 0; i < N; ++i) data[i] = static_cast<T>(p[i]);
  }
};
}  // namespace flashinfer
