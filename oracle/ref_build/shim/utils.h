// TEST INFRASTRUCTURE ONLY: shadows csrc/kernels/utils.h for the host build (its device assert is `asm("trap;")`).
#pragma once
#define CHECK_CUDA(call) (void)(call)
#define CHECK_CUDA_KERNEL_LAUNCH()
#define FLASH_ASSERT(cond) do { if (!(cond)) { fprintf(stderr, "Assertion failed (%s:%d): %s\n", __FILE__, __LINE__, #cond); abort(); } } while (0)
#define FLASH_DEVICE_ASSERT(cond) FLASH_ASSERT(cond)
