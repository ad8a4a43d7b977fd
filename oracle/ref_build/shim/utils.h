// TEST INFRASTRUCTURE ONLY.  Host build of the reference kernels (build_ref.py): this header is found first on the
// include path under the name the reference's kernels include for their error macros, because the original's device
// assert emits a PTX trap instruction that an x86 assembler rejects.  Launch checks have nothing to check on the host;
// a failed device assert aborts the test process.
#pragma once
#include <cstdio>
#include <cstdlib>

namespace refshim {
[[noreturn]] inline void device_assert_failed(const char* what, const char* file, int line) {
  std::fprintf(stderr, "reference kernel assertion (host build) %s at %s:%d\n", what, file, line);
  std::abort();
}
}  // namespace refshim

#define FLASH_DEVICE_ASSERT(cond) ((cond) ? (void)0 : refshim::device_assert_failed(#cond, __FILE__, __LINE__))
#define FLASH_ASSERT(cond) FLASH_DEVICE_ASSERT(cond)
#define CHECK_CUDA(call) ((void)(call))
#define CHECK_CUDA_KERNEL_LAUNCH() ((void)0)
