// UMMA descriptor self-test ("lab"): runs one tcgen05.mma chain per operand-layout variant used by the MLA
// decode and FP8 GEMM kernels and checks it against a host fp32 reference.  It exists because the smem /
// instruction descriptor encodings are hand-built (no CUTLASS); each kernel's layout assumptions are pinned
// here on real hardware before the big kernels rely on them.
//
// Variants (see b200_umma_selftest):
//   0  fp8  A K-major SW128 (TMA)  x  B K-major SW128 (TMA)            -> q.k^T nope term, GEMM main loop
//   1  bf16 A K-major SW128 (TMA)  x  B K-major SW128 (TMA)            -> q.k^T rope term
//   2  fp8  A MN-major SW128 (TMA, the K-tile re-read as V^T)  x  B MN-major no-swizzle (thread-written P)
//   3  as 2 with the LBO/SBO roles of B exchanged (diagnostic)
//   4  fp8  A MN-major SW128 x B K-major SW128 thread-written (fallback layout for P)
//   5  bf16 A MN-major SW128, M = 128 dims = two 64-dim chunks (LBO = chunk stride)  x  B K-major SW128 thread-written
//      bf16 P[q][tok], two 64-token chunks                              -> P.V of the bf16-KV kernel (mla_decode_bf16.cu)
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "b200_lab.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace b200 {

struct LabParams {
  CUtensorMap tm_a, tm_b;
  const uint8_t* b_raw;  // logical [K][N] bytes for thread-written B
  float* out;            // [128][N]
  int n;
  int kind;  // 0 = f8f6f4 (e4m3), 1 = f16 (bf16)
  int a_mn_major, b_mn_major;
  int num_k_steps;
  int a_rows;  // TMA box rows for A
  int b_mode;  // 0 = TMA K-major SW128; 1 = thread-written MN-major no-swizzle; 2 = thread-written K-major SW128
  uint32_t a_layout, b_layout;
  uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
  uint32_t a_kstep, b_kstep;  // bytes added to the start address per k-step
  uint32_t b_chunk_stride, b_group_stride;
  uint32_t k_total;  // K extent (tokens) for thread-written B
  int a_chunks;      // > 1: A is loaded as a_chunks boxes of 64 bf16 columns, 16 KB apart
  uint32_t b_kstep_hi;  // != 0: B start address advances b_kstep per k-step within groups of 4 and b_kstep_hi per group
};

__global__ void __launch_bounds__(128, 1) umma_lab_kernel(const __grid_constant__ LabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;           // 32 KB
  uint8_t* s_b = smem + 32768;   // 32 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;

  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(&tmem_base_slot, 256);  // a fully converged warp must execute the .sync.aligned alloc
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;

  if (tid == 0) {
    uint32_t bytes = p.a_rows * 128 * (p.a_chunks > 1 ? p.a_chunks : 1);
    tma_load_2d(s_a, &p.tm_a, &bar_load, 0, 0);
    for (int c = 1; c < p.a_chunks; ++c) tma_load_2d(s_a + c * 16384, &p.tm_a, &bar_load, c * 64, 0);
    if (p.b_mode == 0) {
      tma_load_2d(s_b, &p.tm_b, &bar_load, 0, 0);
      bytes += p.n * 128;
    }
    mbar_arrive_expect_tx(&bar_load, bytes);
  }
  if (p.b_mode == 1) {
    // thread t = token (K index); writes its N bytes as 16-B chunks: core matrix (group g = t/8, chunk c)
    const int t = tid;
    if (t < (int)p.k_total) {
      for (int c = 0; c < p.n / 16; ++c) {
        uint4 v = *reinterpret_cast<const uint4*>(p.b_raw + (size_t)t * p.n + c * 16);
        *reinterpret_cast<uint4*>(s_b + c * p.b_chunk_stride + (t >> 3) * p.b_group_stride + (t & 7) * 16) = v;
      }
    }
    fence_proxy_async_smem();
  } else if (p.b_mode == 2) {
    // K-major SW128: row n holds K (tokens) contiguously; 16-B chunk index XOR (n & 7)
    for (int idx = tid; idx < p.n * (int)p.k_total; idx += 128) {
      int n = idx / p.k_total, k = idx % p.k_total;
      uint32_t off = n * 128 + ((((k >> 4) ^ (n & 7)) & 7) << 4) + (k & 15);
      s_b[off] = p.b_raw[(size_t)k * p.n + n];
    }
    fence_proxy_async_smem();
  } else if (p.b_mode == 3) {
    // bf16 P[q][tok], K-major SW128, 64-token chunks n * 128 B apart: thread t = token writes its n values
    const uint16_t* raw = reinterpret_cast<const uint16_t*>(p.b_raw);  // logical [K tokens][N]
    const int t = tid;
    if (t < (int)p.k_total) {
      for (int q = 0; q < p.n; ++q) {
        const uint32_t off = (t >> 6) * p.n * 128 + (q >> 3) * 1024 + (q & 7) * 128 +
                             (((((t & 63) >> 3) ^ (q & 7)) & 7) << 4) + (t & 7) * 2;
        *reinterpret_cast<uint16_t*>(s_b + off) = raw[(size_t)t * p.n + q];
      }
    }
    fence_proxy_async_smem();
  }
  __syncthreads();

  if (tid == 0) {
    mbar_wait(&bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc(p.kind ? FMT_BF16 : FMT_E4M3, p.kind ? FMT_BF16 : FMT_E4M3, p.a_mn_major,
                                      p.b_mn_major, 128, p.n);
    for (int k = 0; k < p.num_k_steps; ++k) {
      uint64_t da = make_smem_desc(smem_u32(s_a) + k * p.a_kstep, p.a_lbo, p.a_sbo, p.a_layout);
      const uint32_t b_off = p.b_kstep_hi ? (k & 3) * p.b_kstep + (k >> 2) * p.b_kstep_hi : k * p.b_kstep;
      uint64_t db = make_smem_desc(smem_u32(s_b) + b_off, p.b_lbo, p.b_sbo, p.b_layout);
      if (p.kind) umma_f16(tmem, da, db, idesc, k > 0);
      else umma_f8(tmem, da, db, idesc, k > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  {
    const uint32_t lane_base = (warp & 3) * 32;
    for (int c0 = 0; c0 < p.n; c0 += 16) {
      float v[16];
      tmem_ld<16>(tmem + (lane_base << 16) + c0, v);
      tmem_wait_ld();
      for (int j = 0; j < 16; ++j) p.out[(size_t)(lane_base + (tid & 31)) * p.n + c0 + j] = v[j];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// Issue-rate probe: `iters` back-to-back MMAs of one operand-layout variant on resident (uninitialised) smem.
//   variant 0: fp8 A K-major SW128 x B K-major SW128         (Q.K^T nope, GEMM)
//   variant 1: bf16 A K-major SW128 x B K-major SW128        (Q.K^T rope)
//   variant 2: fp8 A MN-major SW128 x B MN-major no-swizzle  (P.V as shipped)
//   variant 3: fp8 A MN-major SW128 x B K-major SW128        (P.V with a K-major P)
//   variant 4: fp8 A K-major SW128 x B MN-major no-swizzle
__global__ void __launch_bounds__(128, 1) umma_rate_kernel(int variant, int n, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (64 + 32) * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  // Warp 0 issues from warp-uniform code: every lane runs the loop, only the MMA is predicated with elect.sync, so the
  // descriptors stay in uniform registers (variant >= 8 selects the old divergent `if (tid == 0)` issue for contrast).
  const bool divergent = variant >= 8;
  variant &= 7;
  if (warp == 0 && (!divergent || tid == 0)) {
    const uint32_t a = smem_u32(smem), b = a + 64 * 1024;
    const bool bf = variant == 1;
    const uint32_t a_mn = (variant == 2 || variant == 3), b_mn = (variant == 2 || variant == 4);
    const uint32_t idesc = make_idesc(bf ? FMT_BF16 : FMT_E4M3, bf ? FMT_BF16 : FMT_E4M3, a_mn, b_mn, 128, n);
    const uint64_t da0 = a_mn ? make_smem_desc(0, 16384, 1024, SWZ_128B) : make_smem_desc(0, 0, 1024, SWZ_128B);
    const uint64_t db0 = b_mn ? make_smem_desc(0, 128, 2048, SWZ_NONE) : make_smem_desc(0, 0, 1024, SWZ_128B);
    const long long t0 = clock64();
    if (divergent) {
      for (int i = 0; i < iters; ++i) {
        const uint32_t ka = a_mn ? (i & 3) * 4096 + ((i >> 2) & 3) * 16384 : (i & 3) * 32 + ((i >> 2) & 3) * 16384;
        const uint32_t kb = b_mn ? (i & 3) * 512 : (i & 3) * 32;
        const uint64_t da = da0 | (uint64_t)(((a + ka) >> 4) & 0x3FFF);
        const uint64_t db = db0 | (uint64_t)(((b + kb) >> 4) & 0x3FFF);
        if (bf) umma_f16(tmem + (i & 1) * 256, da, db, idesc, 1);
        else umma_f8(tmem + (i & 1) * 256, da, db, idesc, 1);
      }
      umma_commit(&bar);
    } else {
      for (int i = 0; i < iters; i += 16) {
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const uint32_t ka = a_mn ? (j & 3) * 4096 + ((j >> 2) & 3) * 16384 : (j & 3) * 32 + ((j >> 2) & 3) * 16384;
            const uint32_t kb = b_mn ? (j & 3) * 512 : (j & 3) * 32;
            const uint64_t da = da0 | (uint64_t)(((a + ka) >> 4) & 0x3FFF);
            const uint64_t db = db0 | (uint64_t)(((b + kb) >> 4) & 0x3FFF);
            if (bf) umma_f16(tmem + (j & 1) * 256, da, db, idesc, 1);
            else umma_f8(tmem + (j & 1) * 256, da, db, idesc, 1);
          }
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&bar);
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    if (lane_id() == 0) cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// CTA-pair lab: one cta_group::2 MMA chain (M = 128 = 64 rows per CTA, N = n, K = 128 bytes) with operands staged by
// cta_group::2 TMA loads that credit the leader's barrier, completion multicast to both CTAs, and a raw dump of each
// CTA's TMEM (128 lanes x 256 columns).  The host side (tools/first_light.py pair) derives the D layout from the dump.
//   b_mode 0: B K-major   SW128, global [n rows][128 B]; CTA r stages rows [r*n/2, (r+1)*n/2)          (Q.K^T form)
//   b_mode 1: B MN-major  SW128, global [128 k rows][n B] (n = 256); CTA r stages bytes [128 r, 128 r + 128) of every
//             row                                                                                        (P.V form)
// ------------------------------------------------------------------------------------------------------------------
struct PairLabParams {
  CUtensorMap tm_a, tm_b;
  float* out;  // [2][128][256]
  int n, b_mode;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) umma_pair_lab_kernel(const __grid_constant__ PairLabParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_a = smem;          // 64 rows x 128 B = 8 KB
  uint8_t* s_b = smem + 8192;   // <= 16 KB
  __shared__ uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) {
    mbar_init(&bar_load, 2);  // one expect_tx arrival per CTA
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(&tmem_base_slot, 256);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_base_slot;
  {  // clear this CTA's TMEM window so untouched cells read back as a sentinel
    float z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = -12345.f;
    for (int c = 0; c < 256; c += 16) tmem_st<16>(tmem + ((uint32_t)(warp * 32) << 16) + c, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();

  const uint32_t leader_load = mapa_u32(smem_u32(&bar_load), 0);
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t b_bytes = p.b_mode == 0 ? (uint32_t)(p.n / 2) * 128u : 128u * 128u;
      mbar_arrive_expect_tx_cluster(leader_load, 8192u + b_bytes);
      tma_load_2d_pair(s_a, &p.tm_a, leader_load, 0, 64 * (int)rank);
      if (p.b_mode == 0) tma_load_2d_pair(s_b, &p.tm_b, leader_load, 0, (p.n / 2) * (int)rank);
      else tma_load_2d_pair(s_b, &p.tm_b, leader_load, 128 * (int)rank, 0);
    }
    __syncwarp();
    if (rank == 0) {
      mbar_wait_cluster(&bar_load, 0);
      tc_fence_after();
      const uint32_t idesc = make_idesc(FMT_E4M3, FMT_E4M3, 0, p.b_mode == 1, 128, p.n);
      const uint64_t da0 = make_smem_desc(0, 0, 1024, SWZ_128B);
      const uint64_t db0 = p.b_mode == 1 ? make_smem_desc(0, 16384, 1024, SWZ_128B) : make_smem_desc(0, 0, 1024, SWZ_128B);
      const uint32_t a = smem_u32(s_a), b = smem_u32(s_b);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t kb = p.b_mode == 1 ? k * 4096 : k * 32;
          umma_f8_pair(tmem, da0 | (uint64_t)(((a + k * 32) >> 4) & 0x3FFF), db0 | (uint64_t)(((b + kb) >> 4) & 0x3FFF), idesc,
                       k > 0);
        }
        umma_commit_pair(&bar_mma, 3);
      }
      __syncwarp();
    }
  }
  mbar_wait_cluster(&bar_mma, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 256; c0 += 16) {
    float v[16];
    tmem_ld<16>(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_wait_ld();
    for (int j = 0; j < 16; ++j) p.out[((size_t)rank * 128 + tid) * 256 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem, 256);
}


// ------------------------------------------------------------------------------------------------------------------
// CTA-pair issue-rate probe: `iters` back-to-back cta_group::2 MMAs (M = 128 over the pair) on resident smem.
//   mode 0: fp8  A K-major SW128 x B K-major SW128, N = n              (pair Q.K^T nope)
//   mode 1: fp8  A K-major SW128 x B MN-major SW128, N = n (256)       (pair P.V)
//   mode 2: bf16 A K-major SW128 x B K-major SW128, N = n              (pair Q.K^T rope)
//   mode 3: the MLA tile mix: 16 x mode 0 (N=128) + 4 x mode 2 (N=128) + 8 x mode 1 (N=256); cycles per 28-MMA tile / 28
// ------------------------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    umma_pair_rate_kernel(int mode, int n, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  for (int i = tid; i < 128 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x38383838u;
  fence_proxy_async_smem();
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(&slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 0 && rank == 0) {
    const uint32_t a = smem_u32(smem), b = a + 40 * 1024;
    const uint32_t id_qk8 = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, mode == 3 ? 128 : n);
    const uint32_t id_qk16 = make_idesc(FMT_BF16, FMT_BF16, 0, 0, 128, mode == 3 ? 128 : n);
    const uint32_t id_pv = make_idesc(FMT_E4M3, FMT_E4M3, 0, 1, 128, mode == 3 ? 256 : n);
    const uint64_t dk = make_smem_desc(0, 0, 1024, SWZ_128B);
    const uint64_t dv = make_smem_desc(0, 16384, 1024, SWZ_128B);
    auto wa = [](uint64_t d, uint32_t x) { return d | (uint64_t)((x >> 4) & 0x3FFF); };
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (elect_one()) {
        if (mode == 0 || mode == 3) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            umma_f8_pair(tmem + (i & 1) * 64, wa(dk, a + (j >> 2) * 8192 + (j & 3) * 32), wa(dk, b + (j >> 2) * 8192 + (j & 3) * 32),
                         id_qk8, 1);
        }
        if (mode == 2 || mode == 3) {
#pragma unroll
          for (int j = 0; j < (16 >> (mode == 3 ? 2 : 0)); ++j)
            umma_f16_pair(tmem + (i & 1) * 64, wa(dk, a + 32768 + (j & 3) * 32), wa(dk, b + 32768 + (j & 3) * 32), id_qk16, 1);
        }
        if (mode == 1 || mode == 3) {
#pragma unroll
          for (int j = 0; j < (mode == 3 ? 8 : 16); ++j)
            umma_f8_pair(tmem + 128 + (j & 1) * 128, wa(dk, a + (j & 3) * 32), wa(dv, b + 40960 + (j & 1) * 16384 + ((j >> 1) & 3) * 4096),
                         id_pv, 1);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_pair(&bar, 1);
    __syncwarp();
    mbar_wait(&bar, 0);
    if (lane_id() == 0) cycles[0] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) tmem_dealloc_pair(tmem, 512);
}

static float e4m3_to_float(uint8_t c) {
  int s = c >> 7, e = (c >> 3) & 0xF, m = c & 7;
  float v;
  if (e == 0xF && m == 7) v = NAN;
  else if (e == 0) v = ldexpf((float)m / 8.f, -6);
  else v = ldexpf(1.f + (float)m / 8.f, e - 7);
  return s ? -v : v;
}
static float bf16_to_float(uint16_t b) {
  uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

}  // namespace b200

using namespace b200;

// Runs one variant; writes max |got - expected| / max|expected| to *rel_err.  Returns 0 on success.
extern "C" int b200_umma_selftest(int variant, int n, float* rel_err) {
  B200_CHECK(n % 16 == 0 && n >= 16 && n <= 256, "selftest: n must be a multiple of 16 in [16,256]");
  LabParams p;
  memset(&p, 0, sizeof(p));
  p.n = n;
  const int K_BYTES = 128;
  std::vector<uint8_t> hA(128 * K_BYTES * 2), hB((size_t)256 * K_BYTES);  // A: room for the two-chunk bf16 image
  srand(1234 + variant);
  auto rnd_e4m3 = [] {
    uint8_t c;
    do c = rand() & 0xFF;
    while ((c & 0x7F) >= 0x58);  // |v| < 16, no NaN
    return c;
  };
  std::vector<float> expect((size_t)128 * n, 0.f);
  uint8_t *dA = nullptr, *dB = nullptr;
  float* dOut = nullptr;
  B200_CUDA(cudaMalloc(&dA, hA.size()));
  B200_CUDA(cudaMalloc(&dB, hB.size()));
  B200_CUDA(cudaMalloc(&dOut, sizeof(float) * 128 * n));
  B200_CUDA(cudaMemset(dOut, 0xFF, sizeof(float) * 128 * n));
  p.out = dOut;
  int rc = 0;
  if (variant == 0) {
    for (auto& x : hA) x = rnd_e4m3();
    for (auto& x : hB) x = rnd_e4m3();
    for (int m = 0; m < 128; ++m)
      for (int j = 0; j < n; ++j) {
        float acc = 0;
        for (int k = 0; k < 128; ++k) acc += e4m3_to_float(hA[m * 128 + k]) * e4m3_to_float(hB[j * 128 + k]);
        expect[(size_t)m * n + j] = acc;
      }
    p.kind = 0; p.num_k_steps = 4; p.a_rows = 128; p.b_mode = 0;
    p.a_layout = p.b_layout = SWZ_128B; p.a_sbo = p.b_sbo = 1024; p.a_kstep = p.b_kstep = 32;
    rc |= make_tma_2d(&p.tm_a, dA, 1, 128, 128, 128, 128, 128, TmaSwizzle::B128);
    rc |= make_tma_2d(&p.tm_b, dB, 1, 128, n, 128, 128, n, TmaSwizzle::B128);
  } else if (variant == 1) {
    uint16_t* a16 = reinterpret_cast<uint16_t*>(hA.data());
    uint16_t* b16 = reinterpret_cast<uint16_t*>(hB.data());
    for (int i = 0; i < 128 * 64; ++i) a16[i] = (uint16_t)(0x3C00 + (rand() & 0x3FF) + ((rand() & 1) << 15));
    for (int i = 0; i < 256 * 64; ++i) b16[i] = (uint16_t)(0x3C00 + (rand() & 0x3FF) + ((rand() & 1) << 15));
    for (int m = 0; m < 128; ++m)
      for (int j = 0; j < n; ++j) {
        float acc = 0;
        for (int k = 0; k < 64; ++k) acc += bf16_to_float(a16[m * 64 + k]) * bf16_to_float(b16[j * 64 + k]);
        expect[(size_t)m * n + j] = acc;
      }
    p.kind = 1; p.num_k_steps = 4; p.a_rows = 128; p.b_mode = 0;
    p.a_layout = p.b_layout = SWZ_128B; p.a_sbo = p.b_sbo = 1024; p.a_kstep = p.b_kstep = 32;
    rc |= make_tma_2d(&p.tm_a, dA, 2, 64, 128, 128, 64, 128, TmaSwizzle::B128);
    rc |= make_tma_2d(&p.tm_b, dB, 2, 64, n, 128, 64, n, TmaSwizzle::B128);
  } else if (variant >= 2 && variant <= 4) {
    // A image = KV chunk [128 tokens][128 d bytes]; used MN-major: D[d][q] = sum_t A[t][d] * P[t][q]
    for (auto& x : hA) x = rnd_e4m3();
    for (size_t i = 0; i < (size_t)128 * n; ++i) hB[i] = rnd_e4m3();  // P logical [token][q]
    for (int d = 0; d < 128; ++d)
      for (int j = 0; j < n; ++j) {
        float acc = 0;
        for (int t = 0; t < 128; ++t) acc += e4m3_to_float(hA[t * 128 + d]) * e4m3_to_float(hB[(size_t)t * n + j]);
        expect[(size_t)d * n + j] = acc;
      }
    p.kind = 0; p.num_k_steps = 4; p.a_rows = 128; p.k_total = 128;
    p.a_mn_major = 1; p.a_layout = SWZ_128B; p.a_sbo = 1024; p.a_lbo = 16384; p.a_kstep = 32 * 128;
    rc |= make_tma_2d(&p.tm_a, dA, 1, 128, 128, 128, 128, 128, TmaSwizzle::B128);
    p.b_raw = dB;
    if (variant == 4) {
      p.b_mode = 2; p.b_mn_major = 0; p.b_layout = SWZ_128B; p.b_sbo = 1024; p.b_lbo = 0; p.b_kstep = 32;
    } else {
      p.b_mode = 1; p.b_mn_major = 1; p.b_layout = SWZ_NONE;
      p.b_group_stride = 128;   // consecutive 8-token groups (K direction)
      p.b_chunk_stride = 2048;  // next 16-q chunk (N direction): 16 groups * 128 B
      p.b_kstep = 4 * 128;      // 32 tokens
      if (variant == 2) { p.b_lbo = p.b_chunk_stride; p.b_sbo = p.b_group_stride; }
      else              { p.b_lbo = p.b_group_stride; p.b_sbo = p.b_chunk_stride; }
    }
  } else if (variant == 5) {
    // A image = bf16 KV tile [128 tokens][128 dims]; used MN-major: D[d][q] = sum_t A[t][d] * P[t][q]
    B200_CHECK(n <= 64, "selftest variant 5: n <= 64");
    uint16_t* a16 = reinterpret_cast<uint16_t*>(hA.data());
    uint16_t* b16 = reinterpret_cast<uint16_t*>(hB.data());
    for (int i = 0; i < 128 * 128; ++i) a16[i] = (uint16_t)(0x3C00 + (rand() & 0x3FF) + ((rand() & 1) << 15));
    for (int i = 0; i < 128 * n; ++i) b16[i] = (uint16_t)(0x3C00 + (rand() & 0x3FF) + ((rand() & 1) << 15));
    for (int d = 0; d < 128; ++d)
      for (int j = 0; j < n; ++j) {
        float acc = 0;
        for (int t = 0; t < 128; ++t) acc += bf16_to_float(a16[t * 128 + d]) * bf16_to_float(b16[(size_t)t * n + j]);
        expect[(size_t)d * n + j] = acc;
      }
    p.kind = 1; p.num_k_steps = 8; p.a_rows = 128; p.k_total = 128; p.a_chunks = 2;
    p.a_mn_major = 1; p.a_layout = SWZ_128B; p.a_sbo = 1024; p.a_lbo = 16384; p.a_kstep = 16 * 128;
    rc |= make_tma_2d(&p.tm_a, dA, 2, 128, 128, 256, 64, 128, TmaSwizzle::B128);
    p.b_raw = dB;
    p.b_mode = 3; p.b_mn_major = 0; p.b_layout = SWZ_128B; p.b_sbo = 1024; p.b_lbo = 0; p.b_kstep = 32;
    p.b_kstep_hi = n * 128;
  } else {
    set_last_error("selftest: unknown variant %d", variant);
    rc = 1;
  }
  if (rc == 0) {
    cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice);
    const int smem_bytes = 1024 + 32768 + 32768;
    cudaFuncSetAttribute(umma_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    umma_lab_kernel<<<1, 128, smem_bytes>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_last_error("selftest variant %d: kernel failed: %s", variant, cudaGetErrorString(e));
      rc = 2;
    } else {
      std::vector<float> got((size_t)128 * n);
      cudaMemcpy(got.data(), dOut, sizeof(float) * got.size(), cudaMemcpyDeviceToHost);
      float max_ref = 0, max_diff = 0;
      for (size_t i = 0; i < got.size(); ++i) {
        max_ref = fmaxf(max_ref, fabsf(expect[i]));
        float d = fabsf(got[i] - expect[i]);
        if (!(d == d)) d = INFINITY;
        max_diff = fmaxf(max_diff, d);
      }
      *rel_err = max_diff / fmaxf(max_ref, 1e-20f);
    }
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dOut);
  return rc;
}

// Average SM cycles per tcgen05.mma (M=128, N=n, K=32 bytes) for one operand-layout variant; see umma_rate_kernel.
extern "C" int b200_umma_rate(int variant, int n, int iters, float* cycles_per_mma) {
  B200_CHECK(n % 16 == 0 && n >= 16 && n <= 256 && iters > 0 && iters % 16 == 0, "umma_rate: bad arguments");
  long long* d = nullptr;
  B200_CUDA(cudaMalloc(&d, sizeof(long long)));
  const int smem_bytes = 1024 + (64 + 32) * 1024;
  cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  umma_rate_kernel<<<1, 128, smem_bytes>>>(variant, n, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h = 0;
  if (e == cudaSuccess) cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) {
    set_last_error("umma_rate: %s", cudaGetErrorString(e));
    return 2;
  }
  *cycles_per_mma = (float)h / iters;
  return 0;
}

// CTA-pair lab (see umma_pair_lab_kernel).  a: host [128][128] e4m3 bytes; b: host, b_mode 0 -> [n][128], b_mode 1 ->
// [128][n] (n must be 256); out: host float [2][128][256] raw TMEM dump of the two CTAs.
extern "C" int b200_umma_pair_lab(int b_mode, int n, const void* a, const void* b, float* out) {
  B200_CHECK(n % 32 == 0 && n >= 32 && n <= 256 && (b_mode == 0 || (b_mode == 1 && n == 256)), "pair_lab: bad arguments");
  uint8_t *dA = nullptr, *dB = nullptr;
  float* dOut = nullptr;
  B200_CUDA(cudaMalloc(&dA, 128 * 128));
  B200_CUDA(cudaMalloc(&dB, 256 * 128));
  B200_CUDA(cudaMalloc(&dOut, sizeof(float) * 2 * 128 * 256));
  cudaMemcpy(dA, a, 128 * 128, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, b, (size_t)n * 128, cudaMemcpyHostToDevice);
  PairLabParams p;
  memset(&p, 0, sizeof(p));
  p.out = dOut;
  p.n = n;
  p.b_mode = b_mode;
  int rc = make_tma_2d(&p.tm_a, dA, 1, 128, 128, 128, 128, 64, TmaSwizzle::B128);
  if (b_mode == 0) rc |= make_tma_2d(&p.tm_b, dB, 1, 128, n, 128, 128, n / 2, TmaSwizzle::B128);
  else rc |= make_tma_2d(&p.tm_b, dB, 1, 256, 128, 256, 128, 128, TmaSwizzle::B128);
  if (rc == 0) {
    const int smem_bytes = 1024 + 8192 + 16384;
    cudaFuncSetAttribute(umma_pair_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    umma_pair_lab_kernel<<<2, 128, smem_bytes>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      set_last_error("pair_lab: kernel failed: %s", cudaGetErrorString(e));
      rc = 2;
    } else {
      cudaMemcpy(out, dOut, sizeof(float) * 2 * 128 * 256, cudaMemcpyDeviceToHost);
    }
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dOut);
  return rc;
}

// Average SM cycles per cta_group::2 tcgen05.mma for one operand form; see umma_pair_rate_kernel.
extern "C" int b200_umma_pair_rate(int mode, int n, int iters, float* cycles_per_mma) {
  B200_CHECK(mode >= 0 && mode <= 3 && n % 16 == 0 && n >= 32 && n <= 256 && iters > 0, "umma_pair_rate: bad arguments");
  long long* d = nullptr;
  B200_CUDA(cudaMalloc(&d, sizeof(long long)));
  const int smem_bytes = 130 * 1024;
  cudaFuncSetAttribute(umma_pair_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  umma_pair_rate_kernel<<<2, 128, smem_bytes>>>(mode, n, iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long c = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&c, d, sizeof(c), cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) {
    set_last_error("umma_pair_rate: %s", cudaGetErrorString(e));
    return 2;
  }
  const int per_iter = mode == 3 ? 28 : 16;
  *cycles_per_mma = (float)c / ((float)iters * per_iter);
  return 0;
}

extern "C" const char* b200_lab_last_error(void) { return b200::get_last_error(); }
