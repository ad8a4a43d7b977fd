/* Hardware probes behind the kernels' hand-built UMMA descriptors: a separate, test-only library
 * (sglang_fluentllm_b200/_C/libb200lab.so, csrc/umma_lab.cu).  Not part of the drop-in boundary and not loaded by the
 * product: tests/ and tools/ use it to pin descriptor layouts and issue rates on the hardware. */
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* last error of this library (it carries its own copy of the error buffer) */
const char* b200_lab_last_error(void);

/* Hardware self-test of the hand-built UMMA shared-memory / instruction descriptors (csrc/umma_lab.cu).
 * Writes max|got-expected|/max|expected| of one operand-layout variant to *rel_err. */
int b200_umma_selftest(int variant, int n, float* rel_err);

/* Issue-rate probe for the same operand layouts: average SM cycles per tcgen05.mma (M=128, N=n, K=32 B). */
int b200_umma_rate(int variant, int n, int iters, float* cycles_per_mma);

/* CTA-pair (cta_group::2) lab: raw TMEM dump [2][128][256] of one M=128 MMA chain; host pointers (diagnostic). */
int b200_umma_pair_lab(int b_mode, int n, const void* a, const void* b, float* out);

/* Issue-rate probe for cta_group::2 MMAs (M = 128 over a CTA pair): average SM cycles per tcgen05.mma.
 * mode 0/1/2 = Q.K^T fp8 / P.V / Q.K^T bf16 operand forms at N = n, mode 3 = the MLA tile mix (28 MMAs per tile). */
int b200_umma_pair_rate(int mode, int n, int iters, float* cycles_per_mma);

#ifdef __cplusplus
}
#endif
