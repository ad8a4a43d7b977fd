// Block-scaled W8A8 FP8 GEMM for sm_100a:  out[m][n] = bf16( sum_kb sA[kb][i] * sB[n/128][kb] * (A_kb . B_kb^T) ).
//
// Replaces deep_gemm.gemm_fp8_fp8_bf16_nt of the reference (3rdparty/deep_gemm/deep_gemm/jit_kernels/gemm.py:145-226,
// kernels include/deep_gemm/fp8_gemm.cuh:60-399 and the SwapAB variant :409-827, both wgmma / sm_90a only).
// Same two-level accumulation contract: the 128-deep inner sum runs on the tensor core in fp32, the outer sum over
// k-blocks is an fp32 FMA with the product of the activation (1x128) and weight (128x128) scales (fp8_gemm.cuh:298-330).
//
// Machine mapping: weights always sit on the UMMA M side (128 weight rows per CTA = one 128x128 scale block, so the
// weight scale is a CTA-uniform scalar per k-block) and tokens on the UMMA N side (BT = 16..128 columns).  For
// decode (m <= 128) that is the reference's SwapAB layout; for large m it is the same kernel with more token tiles.
//   warp 0: TMA producer (weight tile 128x128 B, token tile BTx128 B, BT activation scales per stage)
//   warp 1: UMMA issuer (4 x kind::f8f6f4 M=128 N=BT K=32 per k-block into one of two TMEM accumulators)
//   warps 2..: promotion + epilogue, 4 lane quadrants x BT/CPT column slices: tcgen05.ld the k-block partial,
//              packed-fp32x2 FMA with the scales, bf16 store.
// Decode shapes stream weights (HBM-bound) but have few weight tiles (9..56): split-K over grid.z spreads the stream
// over all SMs; each split stores its fp32 partial to its own workspace slab with plain coalesced stores and a small
// PDL-chained kernel sums the slabs in split order (deterministic; 17 MB of red.global.add cost 30 us per call).
#include <cuda_bf16.h>
#include <stdlib.h>
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
  float* ws;            // [splits][m][n] fp32 split-K partial slabs
  int m, n, k, splits, kb_per_split;
  // grouped GEMMs (tokens-on-M kernel): 0 plain, 1 contiguous (group_layout = m_indices [m], one group per 128-row
  // block, -1 skips the block), 2 masked (group_layout = masked_m [groups]; A / out are [groups][m_max] rows),
  // 3 offset (group_layout = m_offset [groups + 1]: group g owns rows [m_offset[g], m_offset[g + 1]) of A / out)
  const int32_t* group_layout;
  int group_mode, num_groups, m_max;
  int tiles_y;          // grid.y when it is not ceil(m / 128) (offset layout: an upper bound on the row tiles)
};

template <int BT>
struct GemmSmem {
  static constexpr int kStages = BT <= 32 ? 8 : 6;
  static constexpr int kWBytes = 128 * 128;
  static constexpr int kABytes = BT * 128;
  static constexpr int kABytesPadded = (kABytes + 1023) / 1024 * 1024;
  static constexpr int kStageBytes = kWBytes + kABytesPadded + 1024;  // + scales (<= 512 B), keeps 1024 alignment
  static constexpr int kOffWScale = kStages * kStageBytes + 1024;  // weight scales of this CTA's k range, <= 256 floats
  static constexpr int kTotal = kOffWScale + 1024;
};
static constexpr int kMaxWScaleKb = 256;

template <int BT, int CPT>
__global__ void __launch_bounds__(64 + 128 * (BT / CPT), 1) fp8_gemm_kernel(const __grid_constant__ GemmParams p) {
  using S = GemmSmem<BT>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BT < 32 ? 32 : 2 * BT;
  constexpr int kCS = BT / CPT;
  constexpr int kPromo = 128 * kCS;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  struct Bars {
    uint64_t full[8], empty[8], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
  };
  Bars& bars = *reinterpret_cast<Bars*>(smem + kStages * S::kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * 128, m0 = blockIdx.y * BT;
  const int total_kb = p.k / 128;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int num_kb = min(p.kb_per_split, total_kb - kb0);

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars.full[i], 1);
      mbar_init(&bars.empty[i], 1 + kPromo / 32);  // UMMA commit + one arrival per promotion warp (they read the scales)
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.acc_full[i], 1);
      mbar_init(&bars.acc_empty[i], kPromo / 32);  // one arrival per promotion warp: 512 arrivals on one word serialise
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
  pdl_launch_dependents();  // the kernel behind this one (split-K reduce / next quantiser) waits for this grid to complete anyway

  // Issuer warps run warp-uniform code and predicate only the TMA / MMA instructions with elect.sync (descriptors and
  // coordinates stay in uniform registers; a divergent `if (lane == 0)` costs ~110 cycles per tcgen05.mma).
  if (warp == 0) {
    const uint64_t pol_w = policy_evict_first();  // weights are streamed once per token tile
    const int rows_valid = min(BT, p.m - m0);
    // bulk copies need 16-byte granularity: a_scale_stride is a multiple of 4 floats, so rounding up stays in-row
    const uint32_t scale_bytes = (uint32_t)((rows_valid + 3) / 4 * 4) * 4;
    for (int i = 0; i < num_kb; ++i) {
      const int kb = kb0 + i;
      const int st = i % kStages;
      mbar_wait(&bars.empty[st], ((i / kStages) & 1) ^ 1);
      uint8_t* stage = smem + st * S::kStageBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars.full[st], S::kWBytes + S::kABytes + scale_bytes);
        tma_load_2d_hint(stage, &p.tm_w, &bars.full[st], kb * 128, n0, pol_w);
        tma_load_2d(stage + S::kWBytes, &p.tm_a, &bars.full[st], kb * 128, m0);
        bulk_load_1d(stage + S::kWBytes + S::kABytesPadded, p.a_scales + (int64_t)kb * p.a_scale_stride + m0,
                     scale_bytes, &bars.full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, BT);
    const uint64_t dk = make_smem_desc(0, 0, 1024, SWZ_128B);
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % kStages, buf = i & 1;
      mbar_wait(&bars.acc_empty[buf], ((i >> 1) & 1) ^ 1);
      mbar_wait(&bars.full[st], (i / kStages) & 1);
      tc_fence_after();
      const uint32_t w_addr = smem_u32(smem + st * S::kStageBytes);
      const uint32_t a_addr = w_addr + S::kWBytes;
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_f8(tmem + buf * BT, dk | (uint64_t)(((w_addr + s * 32) >> 4) & 0x3FFF),
                  dk | (uint64_t)(((a_addr + s * 32) >> 4) & 0x3FFF), idesc, s != 0);
        umma_commit(&bars.empty[st]);
        umma_commit(&bars.acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int cs = (warp - 2) >> 2;
    const int cb = cs * CPT;
    const int row = quad * 32 + lane;  // weight row within the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int nb = blockIdx.x;
    // weight scales of this CTA's k range -> shared memory once (a per-k-block __ldg is an L2 round trip on the
    // promotion's critical path: there is no L1 left beside 200 KB of stages)
    float* s_wscale = reinterpret_cast<float*>(smem + S::kOffWScale);
    const bool ws_smem = num_kb <= kMaxWScaleKb;
    if (ws_smem) {
      for (int i = tid - 64; i < num_kb; i += kPromo) s_wscale[i] = __ldg(p.b_scales + (int64_t)nb * total_kb + kb0 + i);
      named_bar_sync(1, kPromo);
    }
    float2 acc[CPT / 2];
#pragma unroll
    for (int j = 0; j < CPT / 2; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % kStages, buf = i & 1;
      const float sb = ws_smem ? s_wscale[i] : __ldg(p.b_scales + (int64_t)nb * total_kb + kb0 + i);
      const float2 sb2 = make_float2(sb, sb);
      mbar_wait(&bars.acc_full[buf], (i >> 1) & 1);  // implies full[st] completed
      tc_fence_after();
      const float* sa = reinterpret_cast<const float*>(smem + st * S::kStageBytes + S::kWBytes + S::kABytesPadded) + cb;
      float v[CPT];
      tmem_ld<CPT>(tmem + buf * BT + cb + lane_addr, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < CPT; j += 4) {
        const float4 s4 = *reinterpret_cast<const float4*>(sa + j);
        acc[j / 2] = fma2(make_float2(v[j], v[j + 1]), mul2(make_float2(s4.x, s4.y), sb2), acc[j / 2]);
        acc[j / 2 + 1] = fma2(make_float2(v[j + 2], v[j + 3]), mul2(make_float2(s4.z, s4.w), sb2), acc[j / 2 + 1]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.acc_empty[buf]);
        mbar_arrive(&bars.empty[st]);
      }
    }
    const int ncol = n0 + row;
    if (p.splits == 1) {
      if (ncol < p.n) {
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (m0 + cb + j < p.m)
            p.out[(int64_t)(m0 + cb + j) * p.n + ncol] = __float2bfloat16_rn(j & 1 ? acc[j / 2].y : acc[j / 2].x);
      }
    } else {
      // split-K: this split's fp32 partial -> its own slab of the workspace (plain coalesced stores: a warp covers 32
      // consecutive weight rows = 128 contiguous bytes per token); splitk_reduce_kernel sums the slabs in split order
      if (ncol < p.n) {
        float* slab = p.ws + (int64_t)blockIdx.z * p.m * p.n;
#pragma unroll
        for (int j = 0; j < CPT; ++j)
          if (m0 + cb + j < p.m) slab[(int64_t)(m0 + cb + j) * p.n + ncol] = j & 1 ? acc[j / 2].y : acc[j / 2].x;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// ------------------------------------------------------------------------------------------------------------------
// m > 64: tokens on the UMMA M side (128 rows), 256 weight rows on the N side.  Each CTA re-reads the activation tile
// once per 256 weight rows (half the L2->SM traffic of the 128-row orientation), the activation scale is a per-thread
// scalar (thread = token row), the weight scale a per-128-column scalar, so promotion is one packed FMA per pair, and
// the epilogue writes 128 contiguous bytes per thread.
// ------------------------------------------------------------------------------------------------------------------
template <int BN>
struct GemmTokMSmem {
  static constexpr int kStages = BN == 256 ? 4 : (BN == 192 ? 5 : (BN == 128 ? 6 : 8));
  static constexpr int kABytes = 128 * 128;   // token tile
  static constexpr int kWBytes = BN * 128;    // weight tile
  static constexpr int kStageBytes = kABytes + kWBytes + 1024;  // + 128 activation scales
  static constexpr int kWBlocks = BN > 128 ? (BN + 127) / 128 : 1;  // 128-row weight scale blocks a tile touches
  static constexpr int kOffWScale = kStages * kStageBytes + 1024;  // [kWBlocks][<= 128 k-blocks] floats
  static constexpr int kTotal = kOffWScale + 1024;
  static_assert(kStages * kStageBytes >= 128 * BN * 4, "the epilogue stages the fp32 tile in the pipeline buffers");
};

// BN = weight rows per CTA (UMMA N): 256 for large m (least activation re-read per weight row), 128 / 64 for decode
// shapes, where narrower tiles fill the SMs without split-K (no fp32 partial round trip, no reduce launch).
//
// KC (split-K in a cluster): the `splits` CTAs that share an output tile form one thread-block cluster along grid.z.
// Each stages its fp32 partial tile in its own shared memory; after a cluster barrier CTA z sums rows z, z + splits, ...
// of all the partials over distributed shared memory, in split order (deterministic, bit-identical to the slab + reduce
// kernel path), and writes bf16.  No fp32 round trip through global memory (a 128 x 256 partial drains at ~28 B/clk per
// SM: ~2.9 us) and no second launch.
template <int BN, bool KC, bool GROUPED = false>
__global__ void __launch_bounds__(64 + 2 * BN, 1) fp8_gemm_tokm_kernel(const __grid_constant__ GemmParams p) {
  using S = GemmTokMSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr int CPT = 64, kPromo = 2 * BN;  // 4 lane quadrants x BN / 64 column slices
  constexpr uint32_t kTmemCols = 2 * BN > 256 ? 512 : 2 * BN;  // (an allocation is a power of two)
  constexpr int kWBlocks = S::kWBlocks;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  struct Bars {
    uint64_t full[kStages], empty[kStages], acc_full[2], acc_empty[2];
    uint32_t tmem_base;
  };
  Bars& bars = *reinterpret_cast<Bars*>(smem + kStages * S::kStageBytes);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.x * BN;
  int m0 = blockIdx.y * 128;           // first A / out row of this tile
  int m_begin = 0;                     // rows before this one are not stored (offset layout: the previous group's)
  int m_end = p.m;                     // rows from here on are not stored
  int a_scale_row = m0;                // row index into the (per group, for masked) activation scale columns
  int w_row0 = n0;                     // first row of the weight tile in the [groups * n] weight matrix
  const float* a_scales = p.a_scales;
  const float* b_scales = p.b_scales;
  if constexpr (GROUPED) {  // its own instantiation: the dense kernel's register allocation is not touched
    // m_grouped_gemm_fp8_fp8_bf16_nt_{contiguous, masked} (deep_gemm/jit_kernels/m_grouped_gemm.py:55-215): the row tile
    // picks its expert's weights; tiles without work leave before anything is allocated
    pdl_wait_prior_grid();  // the layout may come from the kernel right before this one
    int g;
    if (p.group_mode == 1) {
      g = __ldg(p.group_layout + m0);
      if (g < 0 || g >= p.num_groups) return;  // -1: skipped block; anything else out of range is not dereferenced
    } else if (p.group_mode == 3) {
      // m_grouped_gemm_fp8_fp8_bf16_nt_offset (m_grouped_gemm.py:225-310): groups start at arbitrary rows.  A group's
      // tiles start at its first row rounded down to 4 (the activation-scale columns are fetched in 16-byte pieces);
      // rows of a tile outside [m_offset[g], m_offset[g + 1]) are computed and dropped.  blockIdx.y counts tiles over
      // the groups in order; the grid holds an upper bound, surplus CTAs leave.
      int ty = blockIdx.y, lo = 0, hi = 0;
      for (g = 0; g < p.num_groups; ++g) {
        lo = __ldg(p.group_layout + g);
        hi = __ldg(p.group_layout + g + 1);
        const int nt = hi > lo ? (hi - (lo & ~3) + 127) / 128 : 0;
        if (ty < nt) break;
        ty -= nt;
      }
      if (g == p.num_groups || lo < 0 || hi > p.m) return;  // surplus CTA, or offsets outside the row range
      m0 = (lo & ~3) + ty * 128;
      m_begin = lo;
      m_end = hi;
      a_scale_row = m0;
    } else {
      const int tiles_m = (p.m_max + 127) / 128;
      g = blockIdx.y / tiles_m;
      const int ml = (blockIdx.y % tiles_m) * 128, mm = __ldg(p.group_layout + g);
      if (ml >= mm) return;
      m0 = g * p.m_max + ml;
      m_end = g * p.m_max + min(mm, p.m_max);
      a_scale_row = ml;
      a_scales += (int64_t)g * (p.k / 128) * p.a_scale_stride;
    }
    w_row0 = g * p.n + n0;
    b_scales += (int64_t)g * ((p.n + 127) / 128) * (p.k / 128);
  }
  const int total_kb = p.k / 128;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int num_kb = min(p.kb_per_split, total_kb - kb0);

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&bars.full[i], 1);
      mbar_init(&bars.empty[i], 1 + kPromo / 32);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&bars.acc_full[i], 1);
      mbar_init(&bars.acc_empty[i], kPromo / 32);  // one arrival per promotion warp: 512 arrivals on one word serialise
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
  pdl_launch_dependents();  // see fp8_gemm_kernel

  if (warp == 0) {
    const uint64_t pol_w = policy_evict_first();
    const int rows_valid = min(128, m_end - m0);
    const uint32_t scale_bytes = (uint32_t)((rows_valid + 3) / 4 * 4) * 4;
    for (int i = 0; i < num_kb; ++i) {
      const int kb = kb0 + i;
      const int st = i % kStages;
      mbar_wait(&bars.empty[st], ((i / kStages) & 1) ^ 1);
      uint8_t* stage = smem + st * S::kStageBytes;
      if (elect_one()) {
        mbar_arrive_expect_tx(&bars.full[st], S::kABytes + S::kWBytes + scale_bytes);
        tma_load_2d(stage, &p.tm_a, &bars.full[st], kb * 128, m0);
        tma_load_2d_hint(stage + S::kABytes, &p.tm_w, &bars.full[st], kb * 128, w_row0, pol_w);
        bulk_load_1d(stage + S::kABytes + S::kWBytes, a_scales + (int64_t)kb * p.a_scale_stride + a_scale_row, scale_bytes,
                     &bars.full[st]);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, BN);
    const uint64_t dk = make_smem_desc(0, 0, 1024, SWZ_128B);
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % kStages, buf = i & 1;
      mbar_wait(&bars.acc_empty[buf], ((i >> 1) & 1) ^ 1);
      mbar_wait(&bars.full[st], (i / kStages) & 1);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + st * S::kStageBytes);
      const uint32_t w_addr = a_addr + S::kABytes;
      if (elect_one()) {
#pragma unroll
        for (int s = 0; s < 4; ++s)
          umma_f8(tmem + buf * BN, dk | (uint64_t)(((a_addr + s * 32) >> 4) & 0x3FFF),
                  dk | (uint64_t)(((w_addr + s * 32) >> 4) & 0x3FFF), idesc, s != 0);
        umma_commit(&bars.empty[st]);
        umma_commit(&bars.acc_full[buf]);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int cs = (warp - 2) >> 2;             // 0..3: 64-column slice
    const int cb = cs * CPT;
    const int row = quad * 32 + lane;           // token row within the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const int nblk = (n0 + cb) / 128;           // weight scale block of this thread's columns
    const bool blk_valid = (n0 + cb) < p.n;
    // weight scales of this CTA's 128-row block(s) over its k range -> shared memory once (see fp8_gemm_kernel)
    float* s_wscale = reinterpret_cast<float*>(smem + S::kOffWScale);
    const bool ws_smem = kWBlocks * num_kb <= 256;
    if (ws_smem) {
      for (int i = tid - 64; i < kWBlocks * num_kb; i += kPromo) {
        const int blk = n0 / 128 + i / num_kb;
        s_wscale[i] = blk * 128 < p.n ? __ldg(b_scales + (int64_t)blk * total_kb + kb0 + i % num_kb) : 0.f;
      }
      named_bar_sync(1, kPromo);
    }
    const float* my_wscale = s_wscale + (nblk - n0 / 128) * num_kb;  // (a 192-row tile may start in the middle of a block)
    float2 acc[CPT / 2];
#pragma unroll
    for (int j = 0; j < CPT / 2; ++j) acc[j] = make_float2(0.f, 0.f);
    for (int i = 0; i < num_kb; ++i) {
      const int st = i % kStages, buf = i & 1;
      const float sb = ws_smem ? my_wscale[i] : (blk_valid ? __ldg(b_scales + (int64_t)nblk * total_kb + kb0 + i) : 0.f);
      mbar_wait(&bars.acc_full[buf], (i >> 1) & 1);
      tc_fence_after();
      const float sa = reinterpret_cast<const float*>(smem + st * S::kStageBytes + S::kABytes + S::kWBytes)[row];
      const float sc = sa * sb;
      const float2 s2 = make_float2(sc, sc);
      float v[CPT];
      tmem_ld<CPT>(tmem + buf * BN + cb + lane_addr, v);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < CPT / 2; ++j) acc[j] = fma2(make_float2(v[2 * j], v[2 * j + 1]), s2, acc[j]);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&bars.acc_empty[buf]);
        mbar_arrive(&bars.empty[st]);
      }
    }
    // Epilogue through shared memory: a thread owns one token row x 64 columns, so storing from registers scatters
    // 16-byte pieces over 32 rows per instruction (a quarter of the kernel's stall samples at m = 128).  The pipeline
    // stages are free now (every MMA has retired and every promotion thread is past its last scale read after the
    // barrier): stage the tile as [128 rows][256 columns], 16-byte chunks XOR-swizzled with (row & 7), then each warp
    // writes whole rows -- 512 B (bf16) or 2 x 512 B (fp32 split-K partial) of contiguous global memory per instruction.
    named_bar_sync(1, kPromo);
    const int pw = warp - 2;                           // promotion warp
    constexpr int kRowsPerWarp = 128 / (kPromo / 32);  // rows each warp writes out
    constexpr bool kEvenRows = 128 % (kPromo / 32) == 0;
    if (p.splits == 1) {
      constexpr int kChunks = BN / 8;  // 16-byte chunks per bf16 row
      uint8_t* srow = smem + row * (BN * 2);
#pragma unroll
      for (int j = 0; j < CPT; j += 8) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[j / 2].x, acc[j / 2].y);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(acc[j / 2 + 1].x, acc[j / 2 + 1].y);
        __nv_bfloat162 h2 = __floats2bfloat162_rn(acc[j / 2 + 2].x, acc[j / 2 + 2].y);
        __nv_bfloat162 h3 = __floats2bfloat162_rn(acc[j / 2 + 3].x, acc[j / 2 + 3].y);
        *reinterpret_cast<uint4*>(srow + ((((cb + j) >> 3) ^ (row & 7)) << 4)) =
            make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                       *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
      }
      named_bar_sync(1, kPromo);
#pragma unroll
      for (int it = 0; it < 128 * kChunks / kPromo; ++it) {
        // BN = 192: twelve promotion warps do not divide the 128 rows; chunks are dealt round-robin over all of them
        const int idx = kEvenRows ? it * 32 + lane : it * kPromo + (tid - 64);
        const int r = kEvenRows ? pw * kRowsPerWarp + idx / kChunks : idx / kChunks, c = idx % kChunks;
        const uint4 v = *reinterpret_cast<const uint4*>(smem + r * (BN * 2) + ((c ^ (r & 7)) << 4));
        if (m0 + r >= m_begin && m0 + r < m_end && n0 + c * 8 < p.n)
          *reinterpret_cast<uint4*>(p.out + (int64_t)(m0 + r) * p.n + n0 + c * 8) = v;
      }
    } else {
      constexpr int kChunks = BN / 4;  // 16-byte chunks per fp32 row
      uint8_t* srow = smem + row * (BN * 4);
#pragma unroll
      for (int j = 0; j < CPT; j += 4)
        *reinterpret_cast<float4*>(srow + ((((cb + j) >> 2) ^ (row & 7)) << 4)) =
            make_float4(acc[j / 2].x, acc[j / 2].y, acc[j / 2 + 1].x, acc[j / 2 + 1].y);
      if constexpr (!KC) {  // (KC: the staged tile is summed across the cluster below)
        named_bar_sync(1, kPromo);
        float* slab = p.ws + ((int64_t)blockIdx.z * p.m + m0) * p.n + n0;
#pragma unroll
        for (int it = 0; it < 128 * kChunks / kPromo; ++it) {
          const int idx = kEvenRows ? it * 32 + lane : it * kPromo + (tid - 64);
          const int r = kEvenRows ? pw * kRowsPerWarp + idx / kChunks : idx / kChunks, c = idx % kChunks;
          const float4 v = *reinterpret_cast<const float4*>(smem + r * (BN * 4) + ((c ^ (r & 7)) << 4));
          if (m0 + r < p.m && n0 + c * 4 < p.n) *reinterpret_cast<float4*>(slab + (int64_t)r * p.n + c * 4) = v;
        }
      }
    }
  }
  if constexpr (KC) {
    // every partial tile of the cluster is staged ([128 rows][BN] fp32, 16-byte chunks XOR-swizzled with row & 7)
    __syncthreads();
    cluster_sync_all();
    if (warp >= 2) {
      const uint32_t nsplit = (uint32_t)p.splits, zr = cluster_ctarank();
      uint32_t peer[8];
#pragma unroll
      for (uint32_t s = 0; s < 8; ++s) peer[s] = mapa_u32(smem_u32(smem), s < nsplit ? s : 0);
      constexpr int kUnits = BN / 8;  // 8 output columns (two fp32 chunks in, one bf16 chunk out) per thread and step
      for (int u = tid - 64;; u += kPromo) {
        const int r = (int)zr + (u / kUnits) * (int)nsplit, c8 = u % kUnits;
        if (r >= 128) break;
        const uint32_t off0 = r * (BN * 4) + (((2 * c8) ^ (r & 7)) << 4), off1 = r * (BN * 4) + (((2 * c8 + 1) ^ (r & 7)) << 4);
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
        for (uint32_t s = 0; s < 8; ++s) {
          if (s < nsplit) {
            const float4 x = ld_cluster_f32x4(peer[s] + off0), y = ld_cluster_f32x4(peer[s] + off1);
            if (s == 0) a = x, b = y;
            else a.x += x.x, a.y += x.y, a.z += x.z, a.w += x.w, b.x += y.x, b.y += y.y, b.z += y.z, b.w += y.w;
          }
        }
        if (m0 + r < p.m && n0 + c8 * 8 < p.n) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
          *reinterpret_cast<uint4*>(p.out + (int64_t)(m0 + r) * p.n + n0 + c8 * 8) =
              make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                         *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (KC) cluster_sync_all();  // no CTA frees its shared memory while a peer still reads its partial
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// out[m][n] = bf16( sum over splits, in split order, of the fp32 slabs ): deterministic, no atomics.  8 outputs per thread.
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ ws, __nv_bfloat16* __restrict__ out,
                                                            int64_t mn, int splits) {
  pdl_wait_prior_grid();
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i >= mn) return;
  float4 a = __ldcg(reinterpret_cast<const float4*>(ws + i)), b = __ldcg(reinterpret_cast<const float4*>(ws + i + 4));
  // four slabs at a time with every load issued before the first add (one L2 round trip per group instead of one per
  // slab); the sums still run in split order
  for (int s0 = 1; s0 < splits; s0 += 4) {
    float4 c[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (s0 + j < splits) {
        c[j] = __ldcg(reinterpret_cast<const float4*>(ws + (s0 + j) * mn + i));
        d[j] = __ldcg(reinterpret_cast<const float4*>(ws + (s0 + j) * mn + i + 4));
      }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (s0 + j < splits) {
        a.x += c[j].x, a.y += c[j].y, a.z += c[j].z, a.w += c[j].w;
        b.x += d[j].x, b.y += d[j].y, b.z += d[j].z, b.w += d[j].w;
      }
  }
  __nv_bfloat162 h0 = __floats2bfloat162_rn(a.x, a.y), h1 = __floats2bfloat162_rn(a.z, a.w);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(b.x, b.y), h3 = __floats2bfloat162_rn(b.z, b.w);
  *reinterpret_cast<uint4*>(out + i) = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                                                  *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
  pdl_launch_dependents();
}

static int launch_splitk_reduce(const GemmParams& prm, cudaStream_t stream) {
  const int64_t mn = (int64_t)prm.m * prm.n;  // n % 64 == 0: a multiple of 8
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)((mn / 8 + 255) / 256), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;  // its prologue overlaps the GEMM's tail; griddepcontrol.wait orders the reads
  B200_CUDA(cudaLaunchKernelEx(&cfg, splitk_reduce_kernel, (const float*)prm.ws, prm.out, mn, prm.splits));
  return 0;
}

// KC: cluster of prm.splits CTAs along grid.z, reduction over distributed shared memory (2 <= splits <= 8).
// Returns -1 (without launching) when the device cannot keep every cluster of the grid resident at once.
template <int BN, bool KC, bool GROUPED = false>
static int launch_gemm_tokm(const GemmParams& prm, int use_pdl, cudaStream_t stream) {
  const int smem = GemmTokMSmem<BN>::kTotal + 1024;
  // per device and context, so not cached per process (a process may drive several GPUs)
  B200_CUDA(cudaFuncSetAttribute(fp8_gemm_tokm_kernel<BN, KC, GROUPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((prm.n + BN - 1) / BN, prm.tiles_y ? prm.tiles_y : (prm.m + 127) / 128, prm.splits);
  cfg.blockDim = dim3(64 + 2 * BN, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (KC) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 1;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = prm.splits;
    ++na;
    static int max_clusters[9] = {0};  // per cluster size, for this tile shape
    if (max_clusters[prm.splits] == 0) {
      cfg.attrs = attr;
      cfg.numAttrs = na;
      int nc = 0;
      if (cudaOccupancyMaxActiveClusters(&nc, fp8_gemm_tokm_kernel<BN, KC, GROUPED>, &cfg) != cudaSuccess) nc = -1;
      (void)cudaGetLastError();
      max_clusters[prm.splits] = nc > 0 ? nc : -1;
    }
    if (max_clusters[prm.splits] < (int)(cfg.gridDim.x * cfg.gridDim.y)) return -1;
  }
  if (use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  B200_CUDA(cudaLaunchKernelEx(&cfg, fp8_gemm_tokm_kernel<BN, KC, GROUPED>, prm));
  return 0;
}

template <int BT, int CPT>
static int launch_gemm(const GemmParams& prm, int use_pdl, cudaStream_t stream) {
  const int smem = GemmSmem<BT>::kTotal + 1024;
  // per device and context, so not cached per process (a process may drive several GPUs)
  B200_CUDA(cudaFuncSetAttribute(fp8_gemm_kernel<BT, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((prm.n + 127) / 128, (prm.m + BT - 1) / BT, prm.splits);
  cfg.blockDim = dim3(64 + 128 * (BT / CPT), 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, fp8_gemm_kernel<BT, CPT>, prm));
  return 0;
}

static int pick_bt(int m) { return m <= 16 ? 16 : (m <= 32 ? 32 : (m <= 64 ? 64 : 128)); }
static bool tokens_on_m(int m) { return m > 64; }

// split-K factor: fill the SMs when the tile grid is small; keep >= 2 k-blocks per split
static int splits_for(int tiles, int k, int sm_count) {
  const int num_kb = k / 128;
  if (tiles * 2 > sm_count) return 1;
  int splits = sm_count / tiles;
  if (splits > num_kb / 2) splits = num_kb / 2;
  if (splits > 32) splits = 32;
  if (splits < 1) splits = 1;
  const int kb_per = (num_kb + splits - 1) / splits;  // the split count the launch will really use
  return (num_kb + kb_per - 1) / kb_per;
}

struct GemmConfig {
  int bn;      // weight rows per CTA (tokens-on-M kernel) or 128 (weights-on-M kernel)
  int splits;
};

// Tokens-on-M kernel: weight rows per CTA and split-K factor from a per-SM traffic model (bytes through one SM on the
// critical path): waves x k-blocks x (activation tile + weight tile), plus the fp32 partial round trip and the extra
// launch when K is split.  B200_GEMM_BN / B200_GEMM_SPLITS override (tuning).
static GemmConfig pick_config(int m, int n, int k, int sm_count) {
  if (!tokens_on_m(m)) {
    const int bt = pick_bt(m);
    return {128, splits_for(((n + 127) / 128) * ((m + bt - 1) / bt), k, sm_count)};
  }
  static const int force_bn = getenv("B200_GEMM_BN") ? atoi(getenv("B200_GEMM_BN")) : 0;
  static const int force_splits = getenv("B200_GEMM_SPLITS") ? atoi(getenv("B200_GEMM_SPLITS")) : 0;
  const int num_kb = k / 128;
  GemmConfig best = {256, 1};
  double best_cost = 1e30;
  static const int kBn[4] = {256, 192, 128, 64};
  for (int bi = 0; bi < 4; ++bi) {
    const int bn = kBn[bi];
    if (force_bn && bn != force_bn) continue;
    if (bn == 192 && n % 192 != 0) continue;  // 192-row tiles only where they tile n exactly (q_b_proj: 24576 = 128 x 192)
    const int tiles = ((n + bn - 1) / bn) * ((m + 127) / 128);
    int splits = force_splits ? force_splits : splits_for(tiles, k, sm_count);
    if (splits > num_kb) splits = num_kb;
    const int kb_per = (num_kb + splits - 1) / splits;
    splits = (num_kb + kb_per - 1) / kb_per;
    const int waves = (tiles * splits + sm_count - 1) / sm_count;
    double cost = (double)waves * (kb_per * (16384.0 + bn * 128.0) + 128.0 * bn * (splits > 1 ? 4 : 2) * 2.0);
    if (splits > 1) cost += 3.5 * splits * m * n * 4.0 / sm_count + 150e3;  // slab round trip through the reduce kernel + its launch
    if (cost < best_cost) best_cost = cost, best = {bn, splits};
  }
  return best;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_fp8_gemm_workspace_bytes(int m, int n, int k) {
  if (m <= 0 || n <= 0 || k < 128) return 0;
  const int sms = device_sm_count();
  const GemmConfig c = pick_config(m, n, k, sms > 0 ? sms : 148);
  return c.splits == 1 ? 0 : (size_t)c.splits * m * n * sizeof(float);
}

int b200_fp8_gemm_nt(const void* a, const float* a_scales, int64_t a_scale_stride, const void* b, const float* b_scales,
                     void* out, int m, int n, int k, void* workspace, size_t workspace_bytes, int use_pdl,
                     void* stream) {
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
  const int BT = pick_bt(m);
  const int sms = device_sm_count();
  const GemmConfig gc = pick_config(m, n, k, sms > 0 ? sms : 148);
  int splits = gc.splits;
  const size_t need = (size_t)splits * m * n * sizeof(float);
  if (splits > 1 && (workspace == nullptr || workspace_bytes < need)) splits = 1;  // no workspace: plain tiling
  const int num_kb = k / 128;
  prm.splits = splits;
  prm.kb_per_split = (num_kb + splits - 1) / splits;
  prm.splits = (num_kb + prm.kb_per_split - 1) / prm.kb_per_split;
  if (prm.splits > 1) prm.ws = (float*)workspace;
  const bool tokm = tokens_on_m(m);
  if (make_tma_2d(&prm.tm_w, b, 1, (uint64_t)k, (uint64_t)n, (uint64_t)k, 128, tokm ? gc.bn : 128, TmaSwizzle::B128))
    return 1;
  if (make_tma_2d(&prm.tm_a, a, 1, (uint64_t)k, (uint64_t)m, (uint64_t)k, 128, tokm ? 128 : BT, TmaSwizzle::B128))
    return 1;
  prm.a_scales = a_scales;
  prm.a_scale_stride = a_scale_stride;
  prm.b_scales = b_scales;
  prm.out = (__nv_bfloat16*)out;
  prm.m = m;
  prm.n = n;
  prm.k = k;
  int rc = 0;
  bool reduced_in_cluster = false;
  if (tokm) {
    const cudaStream_t st = (cudaStream_t)stream;
    static const bool no_kc = getenv("B200_GEMM_NO_KC") != nullptr;  // A/B knob: always reduce through global slabs
    rc = -1;
    if (prm.splits >= 2 && prm.splits <= 8 && !no_kc) {
      if (gc.bn == 256) rc = launch_gemm_tokm<256, true>(prm, use_pdl, st);
      else if (gc.bn == 192) rc = launch_gemm_tokm<192, true>(prm, use_pdl, st);
      else if (gc.bn == 128) rc = launch_gemm_tokm<128, true>(prm, use_pdl, st);
      else rc = launch_gemm_tokm<64, true>(prm, use_pdl, st);
      reduced_in_cluster = rc == 0;
    }
    if (rc == -1) {  // no split, more than 8 splits, or the clusters do not all fit at once
      if (gc.bn == 256) rc = launch_gemm_tokm<256, false>(prm, use_pdl, st);
      else if (gc.bn == 192) rc = launch_gemm_tokm<192, false>(prm, use_pdl, st);
      else if (gc.bn == 128) rc = launch_gemm_tokm<128, false>(prm, use_pdl, st);
      else rc = launch_gemm_tokm<64, false>(prm, use_pdl, st);
    }
  }
  else if (BT == 16) rc = launch_gemm<16, 8>(prm, use_pdl, (cudaStream_t)stream);
  else if (BT == 32) rc = launch_gemm<32, 8>(prm, use_pdl, (cudaStream_t)stream);
  else rc = launch_gemm<64, 16>(prm, use_pdl, (cudaStream_t)stream);
  if (rc == 0 && prm.splits > 1 && !reduced_in_cluster) rc = launch_splitk_reduce(prm, (cudaStream_t)stream);
  return rc;
}

/* Row tiles (grid.y) a grouped launch carries.  Offset layout: a non-empty group of len rows starting up to 3 rows into
 * its first tile has ceil((len + 3) / 128) <= (len + 130) / 128 tiles, so all groups together have at most
 * floor((m + 130 * groups) / 128); surplus CTAs leave at once. */
int b200_fp8_gemm_grouped_row_tiles(int mode, int num_groups, int m) {
  if (num_groups <= 0 || m < 0) return 0;
  if (mode == 1) return (m + 127) / 128;
  if (mode == 2) return (int)(((int64_t)num_groups * m + 127) / 128);
  if (mode == 3) return (int)(((int64_t)m + 130 * (int64_t)num_groups) / 128);
  return 0;
}

/* Grouped GEMMs on the tokens-on-M kernel: each 128-row tile multiplies its group's weights. */
int b200_fp8_gemm_grouped_nt(const void* a, const float* a_scales, int64_t a_scale_stride, const void* b,
                             const float* b_scales, void* out, const int32_t* group_layout, int mode, int num_groups,
                             int m, int n, int k, int use_pdl, void* stream) {
  B200_CHECK(a && a_scales && b && b_scales && out && group_layout, "fp8_gemm_grouped: null pointer");
  B200_CHECK(mode >= 1 && mode <= 3, "fp8_gemm_grouped: mode must be 1 (contiguous), 2 (masked) or 3 (offset)");
  B200_CHECK(n % 128 == 0 && k % 128 == 0 && n > 0 && k > 0 && num_groups > 0,
             "fp8_gemm_grouped: need n %% 128 == 0, k %% 128 == 0 (n=%d k=%d groups=%d)", n, k, num_groups);
  B200_CHECK(m >= 0, "fp8_gemm_grouped: negative m");
  if (m == 0) return 0;
  // contiguous: m = m_sum, a multiple of the 128-row alignment (get_m_alignment_for_contiguous_layout);
  // masked: m = m_max rows per group; offset: m = m_sum, any value
  B200_CHECK(mode != 1 || m % 128 == 0, "fp8_gemm_grouped: contiguous layout needs m %% 128 == 0 (m=%d)", m);
  B200_CHECK(a_scale_stride >= m && a_scale_stride % 4 == 0,
             "fp8_gemm_grouped: activation scales must be column-major with the row count padded to 4");
  GemmParams prm;
  memset(&prm, 0, sizeof(prm));
  const int bn = n % 256 == 0 ? 256 : 128;
  const int64_t rows = mode == 2 ? (int64_t)num_groups * m : m;
  if (make_tma_2d(&prm.tm_w, b, 1, (uint64_t)k, (uint64_t)num_groups * n, (uint64_t)k, 128, bn, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_a, a, 1, (uint64_t)k, (uint64_t)rows, (uint64_t)k, 128, 128, TmaSwizzle::B128)) return 1;
  prm.a_scales = a_scales;
  prm.a_scale_stride = a_scale_stride;
  prm.b_scales = b_scales;
  prm.out = (__nv_bfloat16*)out;
  prm.m = (int)rows;
  prm.n = n;
  prm.k = k;
  prm.splits = 1;
  prm.kb_per_split = k / 128;
  prm.group_layout = group_layout;
  prm.group_mode = mode;
  prm.num_groups = num_groups;
  prm.m_max = m;
  // launch_gemm_tokm sizes grid.y from prm.m: [groups * m_max] rows only tile per group when m_max is a multiple of 128
  B200_CHECK(mode != 2 || m % 128 == 0 || num_groups == 1,
             "fp8_gemm_grouped: masked layout needs m_max %% 128 == 0 for more than one group (m_max=%d)", m);
  if (mode == 3) prm.tiles_y = b200_fp8_gemm_grouped_row_tiles(mode, num_groups, m);
  return bn == 256 ? launch_gemm_tokm<256, false, true>(prm, use_pdl, (cudaStream_t)stream)
                   : launch_gemm_tokm<128, false, true>(prm, use_pdl, (cudaStream_t)stream);
}

}  // extern "C"
