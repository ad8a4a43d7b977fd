// TEST INFRASTRUCTURE ONLY: the one CUTLASS helper get_mla_metadata.cu uses (cutlass/fast_math.h: ceil_div).
#pragma once
namespace cutlass {
template <class A, class B>
constexpr auto ceil_div(A a, B b) { return (a + b - 1) / b; }
}  // namespace cutlass
