// Block-scaled W8A8 FP8 GEMM for sm_100a:  out[m][n] = bf16( sum_kb sA[kb][i] * sB[n/128][kb] * (A_kb . B_kb^T) ).
//
// Replaces deep_gemm.gemm_fp8_fp8_bf16_nt of the reference (3rdparty/deep_gemm/deep_gemm/jit_kernels/gemm.py:145-226,
// kernels include/deep_gemm/fp8_gemm.cuh:60-399 and the SwapAB variant :409-827, both wgmma / sm_90a only).
// Same two-level accumulation contract: the 128-deep inner sum runs on the tensor core in fp32, the outer sum over
// k-blocks is an fp32 FMA with the product of the activation (1x128) and weight (128x128) scales (fp8_gemm.cuh:298-330).
//
// Machine mapping: weights always sit on the UMMA M side (128 weight rows per CTA = one 128x128 scale block, so the
// weight scale is a CTA-uniform scalar per k-block) and tokens on the UMMA N side (BLOCK_T = 16..128 columns).  For
// decode (m <= 128) that is the reference's SwapAB layout; for large m it is the same kernel with more token tiles.
//   warp 0: TMA producer (weight tile 128x128 B, token tile BLOCK_Tx128 B, BLOCK_T activation scales per stage)
//   warp 1: UMMA issuer (4 x kind::f8f6f4 M=128 N=BLOCK_T K=32 per k-block into one of two TMEM accumulators)
//   warps 2-5: promotion + epilogue (tcgen05.ld the k-block partial, fp32 FMA with the scales, bf16 store)
#include <cuda_bf16.h>
#include <string.h>

#include "b200_decode.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace b200 {

struct GemmParams {
  CUtensorMap tm_w, tm_a;
  const float* a_scales;
  int64_t a_scale_stride;
  const float* b_scales;
  __nv_bfloat16* out;
  int m, n, k;
};

static constexpr int kGemmThreads = 192;

template <int BT>
struct GemmSmem {
  static constexpr int kStages = BT <= 32 ? 8 : 6;
  static constexpr int kWBytes = 128 * 128;
  static constexpr int kABytes = BT * 128;
  static constexpr int kABytesPadded = (kABytes + 1023) / 1024 * 1024;
  static constexpr int kSBytes = 512;  // up to 128 fp32 scales
  static constexpr int kStageBytes = kWBytes + kABytesPadded + kSBytes + 512;  // keep stages 1024-aligned
  static_assert(kStageBytes % 1024 == 0, "stage must stay 1024-byte aligned");
  static constexpr int kTotal = kStages * kStageBytes + 1024;
};

template <int BT>
__global__ void __launch_bounds__(kGemmThreads, 1) fp8_gemm_kernel(const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BT>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BT < 32 ? 32 : 2 * BT;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  struct Bars {
    uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
  };
  Bars& bars = *reinterpret_cast<Bars*>(smem + kStages * S::kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * BT;
  const int num_kb = p.k / 128;

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars.full[i], 1);
      mbar_init(&bars.empty[i], 1 + 128);  // UMMA commit + the 128 promotion threads (they read the scales)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.acc_full[i], 1);
      mbar_init(&bars.acc_empty[i], 128);
    }
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_a);
  }
  if (warp == 1) tmem_alloc(&bars.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  pdl_wait_prior_grid();

  if (warp == 0) {
    if (lane == 0) {
      const uint64_t pol_w = policy_evict_first();  // weights are streamed once per token tile
      const int rows_valid = min(BT, p.m - m0);
      // bulk copies need 16-byte granularity: a_scale_stride is a multiple of 4 floats, so rounding up stays in-row
      const uint32_t scale_bytes = (uint32_t)((rows_valid + 3) / 4 * 4) * 4;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages;
        mbar_wait(&bars.empty[st], ((kb / kStages) & 1) ^ 1);
        uint8_t* stage = smem + st * S::kStageBytes;
        mbar_arrive_expect_tx(&bars.full[st], S::kWBytes + S::kABytes + scale_bytes);
        tma_load_2d_hint(stage, &p.tm_w, &bars.full[st], kb * 128, n0, pol_w);
        tma_load_2d(stage + S::kWBytes, &p.tm_a, &bars.full[st], kb * 128, m0);
        bulk_load_1d(stage + S::kWBytes + S::kABytesPadded, p.a_scales + (int64_t)kb * p.a_scale_stride + m0,
                     scale_bytes, &bars.full[st]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, BT);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int st = kb % kStages, buf = kb & 1;
        mbar_wait(&bars.acc_empty[buf], ((kb >> 1) & 1) ^ 1);
        mbar_wait(&bars.full[st], (kb / kStages) & 1);
        tc_fence_after();
        const uint32_t w_addr = smem_u32(smem + st * S::kStageBytes);
        const uint32_t a_addr = w_addr + S::kWBytes;
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_f8(tmem + buf * BT, make_smem_desc(w_addr + s * 32, 0, 1024, SWZ_128B),
                  make_smem_desc(a_addr + s * 32, 0, 1024, SWZ_128B), idesc, s != 0);
        umma_commit(&bars.empty[st]);
        umma_commit(&bars.acc_full[buf]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;  // weight row within the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int nb = blockIdx.x;
    float acc[BT];
#pragma unroll
    for (int j = 0; j < BT; ++j) acc[j] = 0.f;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % kStages, buf = kb & 1;
      const float sb = __ldg(p.b_scales + (int64_t)nb * num_kb + kb);
      mbar_wait(&bars.acc_full[buf], (kb >> 1) & 1);  // implies full[st] completed
      tc_fence_after();
      const float* sa = reinterpret_cast<const float*>(smem + st * S::kStageBytes + S::kWBytes + S::kABytesPadded);
#pragma unroll
      for (int c0 = 0; c0 < BT; c0 += 16) {
        float v[16];
        tmem_ld<16>(tmem + buf * BT + c0 + lane_addr, v);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 s4 = *reinterpret_cast<const float4*>(sa + c0 + j);
          acc[c0 + j + 0] = fmaf(v[j + 0], s4.x * sb, acc[c0 + j + 0]);
          acc[c0 + j + 1] = fmaf(v[j + 1], s4.y * sb, acc[c0 + j + 1]);
          acc[c0 + j + 2] = fmaf(v[j + 2], s4.z * sb, acc[c0 + j + 2]);
          acc[c0 + j + 3] = fmaf(v[j + 3], s4.w * sb, acc[c0 + j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&bars.acc_empty[buf]);
      mbar_arrive(&bars.empty[st]);
    }
    const int ncol = n0 + row;
    if (ncol < p.n) {
#pragma unroll
      for (int j = 0; j < BT; ++j)
        if (m0 + j < p.m) p.out[(int64_t)(m0 + j) * p.n + ncol] = __float2bfloat16_rn(acc[j]);
    }
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

template <int BT>
static int launch_gemm(const GemmParams& prm, int use_pdl, cudaStream_t stream) {
  const int smem = GemmSmem<BT>::kTotal + 1024;
  static bool configured = false;
  if (!configured) {
    B200_CUDA(cudaFuncSetAttribute(fp8_gemm_kernel<BT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((prm.n + 127) / 128, (prm.m + BT - 1) / BT, 1);
  cfg.blockDim = dim3(kGemmThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, fp8_gemm_kernel<BT>, prm));
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_fp8_gemm_nt(const void* a, const float* a_scales, int64_t a_scale_stride, const void* b,
                                const float* b_scales, void* out, int m, int n, int k, int use_pdl, void* stream) {
  B200_CHECK(a && a_scales && b && b_scales && out, "fp8_gemm: null pointer");
  // same shape contract as the reference: gemm.py:169 (n % 64 == 0, k % 128 == 0)
  B200_CHECK(n % 64 == 0 && k % 128 == 0 && n > 0 && k > 0, "fp8_gemm: need n %% 64 == 0 and k %% 128 == 0 (n=%d k=%d)", n,
             k);
  B200_CHECK(m >= 0, "fp8_gemm: negative m");
  if (m == 0) return 0;  // gemm.py:176-177
  B200_CHECK(a_scale_stride >= m && a_scale_stride % 4 == 0,
             "fp8_gemm: activation scales must be column-major with the row count padded to 4 (stride %lld, m %d)",
             (long long)a_scale_stride, m);
  B200_CHECK((reinterpret_cast<uintptr_t>(a) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(a_scales) & 15) == 0,
             "fp8_gemm: operands must be 16-byte aligned");
  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  const int BT = m <= 16 ? 16 : (m <= 32 ? 32 : (m <= 64 ? 64 : 128));
  if (make_tma_2d(&prm.tm_w, b, 1, (uint64_t)k, (uint64_t)n, (uint64_t)k, 128, 128, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_a, a, 1, (uint64_t)k, (uint64_t)m, (uint64_t)k, 128, BT, TmaSwizzle::B128)) return 1;
  prm.a_scales = a_scales;
  prm.a_scale_stride = a_scale_stride;
  prm.b_scales = b_scales;
  prm.out = (__nv_bfloat16*)out;
  prm.m = m;
  prm.n = n;
  prm.k = k;
  switch (BT) {
    case 16: return launch_gemm<16>(prm, use_pdl, (cudaStream_t)stream);
    case 32: return launch_gemm<32>(prm, use_pdl, (cudaStream_t)stream);
    case 64: return launch_gemm<64>(prm, use_pdl, (cudaStream_t)stream);
    default: return launch_gemm<128>(prm, use_pdl, (cudaStream_t)stream);
  }
}
