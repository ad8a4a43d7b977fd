// Paged MLA decode attention over a bf16 KV cache, sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces the reference's bf16 flash_fwd_splitkv_mla_kernel (3rdparty/flashmla-fp8/csrc/kernels/splitkv_mla.cu,
// host side csrc/flash_api.cpp:65-214; Python flash_mla_interface.py:26-66) -- the kernel the FlashMLA backend
// calls when the KV pool is *not* per-token FP8 (python/sglang/srt/layers/attention/flashmla_backend.py:240-256).
// Same split-KV work partition (tile-scheduler metadata) and the same arithmetic contract: bf16 x bf16 -> f32 logits
// over all 576 dims, base-2 online softmax in f32, P rounded to bf16 before P.V, f32 accumulation, row sums from the
// f32 P (splitkv_mla.cu: the bf16 kernel's softmax), out = O / L -> bf16, lse = ln L + m.
//
// Machine mapping (SwapAB, as mla_decode.cu): KV tokens on the UMMA M side (128 tokens = 2 pages per tile), HB = 32 q
// rows per CTA on the N side; sibling q-block CTAs (grid.x) stream the same pages and meet in L2.
//     S^T[128 tok, HB] = K[128, 576] . Q^T          (36 x kind::f16, K-major SW128 straight from TMA)
//     O^T[512 d,  HB] += V^T[512, 128 tok] . P      (4 dim groups x 8 x kind::f16; V^T = the K tile's first 512 dims
//                                                    described MN-major: two 64-dim chunks per M = 128, LBO = chunk
//                                                    stride; P[q][tok] written K-major SW128 by the softmax threads)
// A 128-token bf16 tile is 144 KB, so the tile is single-buffered and pipelined by column group instead: the producer
// refills each of the five column groups (4 x 128 nope dims, rope) as soon as the MMAs that read it retire.  This is
// the parity row, not the metric path; the FP8 kernels are the tuned ones.
//
// Lazy softmax reference as in mla_decode.cu: bf16 P has the same relative precision at any magnitude, so the
// per-row reference m_ref moves (and O is rescaled in TMEM) only when a logit exceeds it by more than 2^kTrigLog2.
#include <stdlib.h>
#include <string.h>

#include "mla_common.cuh"

namespace b200 {
namespace bf16kv {

static constexpr int HB = 32;                        // q rows per CTA (UMMA N)
static constexpr int CPT = 8;                        // S columns per softmax thread
static constexpr int kCS = HB / CPT;                 // column slices
static constexpr int kSmThreads = 128 * kCS;         // softmax group size
static constexpr int kThreads = 128 + kSmThreads;
static constexpr int kChunkBytes = kTileTokens * 128;          // [128 tok][64 dims] bf16, SW128: 16 KB
static constexpr int kChunks = 9;                              // 8 nope + 1 rope
static constexpr int kGroups = 5;                              // 4 x (2 nope chunks) + rope
static constexpr int kPageChunkBytes = kPage * 128;            // one page of one chunk: 8 KB
static constexpr int kKBytes = kChunks * kChunkBytes;          // 144 KB
static constexpr int kQBytes = kChunks * HB * 128;             // 36 KB
static constexpr int kPBytes = 2 * HB * 128;                   // [2 token chunks][HB rows][128 B]
static constexpr int kOffQ = kKBytes;
static constexpr int kOffP = kOffQ + kQBytes;
static constexpr int kOffSmall = kOffP + kPBytes;
static constexpr int kSmemTotal = kOffSmall + 2048;
static constexpr float kTrigLog2 = 20.f;                       // headroom of the lazy reference
static constexpr uint32_t kTmemCols = 256;                     // S (HB) + O (4 x HB) = 160 -> 256
static constexpr uint32_t kColS = 0, kColO = HB;

struct SmallSmem {
  uint64_t k_full[kGroups], k_empty[kGroups];
  uint64_t q_full, q_empty;
  uint64_t s_full, p_full;
  uint64_t pv_done, o_ready, o_free;
  uint32_t tmem_base;
  uint32_t trig[3];
  float mref[HB];
  float fac[HB];
  float lsum[HB];
  int32_t lim[HB];
  float red[4][HB];
};
static_assert(sizeof(SmallSmem) <= 2048, "SmallSmem too large");

__device__ __forceinline__ uint32_t group_chunk0(int grp) { return grp < 4 ? 2 * grp : 8; }
__device__ __forceinline__ uint32_t group_chunks(int grp) { return grp < 4 ? 2 : 1; }

__global__ void __launch_bounds__(kThreads, 1) mla_decode_bf16_kernel(const __grid_constant__ MlaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_kv = smem;
  uint8_t* s_q = smem + kOffQ;
  uint8_t* s_p = smem + kOffP;
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(smem + kOffSmall);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_block = blockIdx.x, part = blockIdx.z;
  const int q_row0 = q_block * HB;
  const int n_valid_q = min(HB, p.M - q_row0);

  pdl_wait_prior_grid();  // launched as a programmatic dependent of the kernel that produced q / the new K rows
  const int32_t* meta = p.meta + part * B200_MLA_METADATA_INTS;
  const int begin_idx = __ldg(meta + 0), begin_seqlen = __ldg(meta + 1), end_idx = __ldg(meta + 2),
            end_seqlen = __ldg(meta + 3), begin_n_split = __ldg(meta + 4);
  pdl_launch_dependents();  // the combine kernel behind this one orders its reads with griddepcontrol.wait
  if (begin_idx >= p.batch) return;

  if (tid == 0) {
    for (int i = 0; i < kGroups; ++i) {
      mbar_init(&sm.k_full[i], 1);
      mbar_init(&sm.k_empty[i], 1);
    }
    mbar_init(&sm.q_full, 1);
    mbar_init(&sm.q_empty, 1);
    mbar_init(&sm.s_full, 1);
    mbar_init(&sm.p_full, kSmThreads);
    mbar_init(&sm.pv_done, 1);
    mbar_init(&sm.o_ready, 1);
    mbar_init(&sm.o_free, kSmThreads);
    sm.trig[0] = sm.trig[1] = sm.trig[2] = 0;
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_q_nope);
    tma_prefetch_desc(&p.tm_k_nope);
  }
  if (warp == 1) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    uint32_t g = 0, seg_i = 0;
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      if (sg.n_tiles == 0) continue;
      mbar_wait(&sm.q_empty, (seg_i & 1) ^ 1);
      ++seg_i;
      const int qrow = b * p.M + q_row0;
      if (elect_one()) {
        mbar_arrive_expect_tx(&sm.q_full, kQBytes);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) tma_load_2d(s_q + c * HB * 128, &p.tm_q_nope, &sm.q_full, c * 64, qrow);
      }
      __syncwarp();
      const int32_t* bt = p.block_table + (int64_t)b * p.block_table_stride;
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const int blk0 = sg.start_block + 2 * t;
        const int npages = min(2, sg.end_block - blk0);
        const int tok0 = __ldg(bt + blk0) * kPage;
        const int tok1 = npages > 1 ? __ldg(bt + blk0 + 1) * kPage : 0;
        // rope first: its MMAs retire right after Q.K^T of the previous tile, the nope groups after its P.V
#pragma unroll
        for (int gi = 0; gi < kGroups; ++gi) {
          const int grp = (gi + 4) % kGroups;
          mbar_wait(&sm.k_empty[grp], (g & 1) ^ 1);
          if (elect_one()) {
            const uint32_t c0 = group_chunk0(grp), nc = group_chunks(grp);
            mbar_arrive_expect_tx(&sm.k_full[grp], npages * nc * kPageChunkBytes);
            for (uint32_t c = c0; c < c0 + nc; ++c) {
              tma_load_2d(s_kv + c * kChunkBytes, &p.tm_k_nope, &sm.k_full[grp], c * 64, tok0);
              if (npages > 1)
                tma_load_2d(s_kv + c * kChunkBytes + kPageChunkBytes, &p.tm_k_nope, &sm.k_full[grp], c * 64, tok1);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer: S^T = K . Q^T, then O^T += V^T . P ==============================
    constexpr uint32_t idesc_qk = make_idesc(FMT_BF16, FMT_BF16, 0, 0, 128, HB);
    constexpr uint32_t idesc_pv = make_idesc(FMT_BF16, FMT_BF16, 1, 0, 128, HB);
    const uint32_t kv_addr = smem_u32(s_kv), q_addr = smem_u32(s_q), p_addr = smem_u32(s_p);
    const uint64_t dk_sw128 = make_smem_desc(0, 0, 1024, SWZ_128B);            // K-major, 128-B swizzle
    const uint64_t dv_sw128 = make_smem_desc(0, kChunkBytes, 1024, SWZ_128B);  // MN-major view: LBO = next 64 dims
    auto with_addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3FFF); };
    uint32_t g = 0, seg = 0;
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      if (sg.n_tiles == 0) continue;
      mbar_wait(&sm.q_full, seg & 1);
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        // ---- Q.K^T: column groups in landing order (rope, then the four nope groups) ----
#pragma unroll
        for (int gi = 0; gi < kGroups; ++gi) {
          const int grp = (gi + 4) % kGroups;
          mbar_wait(&sm.k_full[grp], g & 1);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t c0 = group_chunk0(grp), nc = group_chunks(grp);
            for (uint32_t c = c0; c < c0 + nc; ++c)
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_f16(tmem + kColS, with_addr(dk_sw128, kv_addr + c * kChunkBytes + s * 32),
                         with_addr(dk_sw128, q_addr + c * HB * 128 + s * 32), idesc_qk, (gi | (int)(c - c0) | s) != 0);
            if (grp == 4) umma_commit(&sm.k_empty[4]);  // (also covers nothing else: first in program order)
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&sm.s_full);
          if (t == sg.n_tiles - 1) umma_commit(&sm.q_empty);
        }
        __syncwarp();
        // ---- P.V ----
        mbar_wait(&sm.p_full, g & 1);
        if (t == 0 && seg > 0) mbar_wait(&sm.o_free, (seg - 1) & 1);
        tc_fence_after();
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 15) / 16;
        if (elect_one()) {
          for (int j = 0; j < 4; ++j) {
            for (int s = 0; s < nk; ++s)
              umma_f16(tmem + kColO + j * HB, with_addr(dv_sw128, kv_addr + 2 * j * kChunkBytes + s * 2048),
                       with_addr(dk_sw128, p_addr + (s >> 2) * HB * 128 + (s & 3) * 32), idesc_pv, (t | s) != 0);
            umma_commit(&sm.k_empty[j]);
          }
          umma_commit(&sm.pv_done);
          if (t == sg.n_tiles - 1) umma_commit(&sm.o_ready);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue ==============================
    const int st_id = tid - 128;
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int cs = (warp - 4) >> 2;             // column slice
    const int cb = cs * CPT;                    // first q column of this thread
    const int tl = quad * 32 + lane;            // token lane within the tile
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const float f = p.scale_log2;
    uint32_t g = 0, o_seg = 0;
    float lt[CPT];                              // per-lane partial row sums relative to mref

    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      named_bar_sync(1, kSmThreads);  // previous request's readers of sm.* are done
      if (st_id < HB) {
        const int row = q_row0 + st_id;
        const bool vq = st_id < n_valid_q;
        sm.mref[st_id] = kMInit;
        sm.lsum[st_id] = 0.f;
        int lim = 0;
        if (vq) lim = p.is_causal ? min(sg.seqlen - mask_len(p.s_q, p.h_q, p.M, row), sg.end_block * kPage) : sg.valid_end;
        sm.lim[st_id] = lim;
      }
      named_bar_sync(1, kSmThreads);
#pragma unroll
      for (int q = 0; q < CPT; ++q) lt[q] = 0.f;

      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        const int token = tile_start + tl;
        const bool tok_valid = token < sg.valid_end;
        if (st_id == 0) sm.trig[(g + 1) % 3] = 0;

        mbar_wait(&sm.s_full, g & 1);
        tc_fence_after();
        float x[CPT];
        const uint32_t s_addr = tmem + kColS + cb + lane_addr;
        tmem_ld<CPT>(s_addr, x);
        tmem_wait_ld();
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          float z = x[q] * f;
          if (!(token < sm.lim[cb + q])) z = -INFINITY;
          x[q] = z;
        }

        bool trigger = (t == 0);
        if (!trigger) {
          float vmax = -INFINITY;
#pragma unroll
          for (int q = 0; q < CPT; ++q) vmax = fmaxf(vmax, x[q] - sm.mref[cb + q]);
          const bool over = vmax > kTrigLog2;
          if (__any_sync(0xffffffffu, over) && lane == 0) sm.trig[g % 3] = 1;
          named_bar_sync(1, kSmThreads);
          trigger = sm.trig[g % 3] != 0;
        }
        if (trigger) {
          // ---- move the reference to the running row maximum, rescale O and L ----
#pragma unroll
          for (int q = 0; q < CPT; ++q) {
            const float wm = warp_max_f32(x[q]);
            if (((cb + q) & 31) == lane) sm.red[quad][cb + q] = wm;
          }
          named_bar_sync(1, kSmThreads);
          if (st_id < HB) {
            const float cm = fmaxf(fmaxf(sm.red[0][st_id], sm.red[1][st_id]), fmaxf(sm.red[2][st_id], sm.red[3][st_id]));
            const float m_old = sm.mref[st_id];
            const float m_new = fmaxf(m_old, cm);
            sm.fac[st_id] = exp2f(m_old - m_new);
            sm.mref[st_id] = m_new;
          }
          named_bar_sync(1, kSmThreads);
          if (t != 0) {
            mbar_wait(&sm.pv_done, (g - 1) & 1);  // every P.V up to the previous tile has landed in O
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              float o[CPT];
              const uint32_t a = tmem + kColO + j * HB + cb + lane_addr;
              tmem_ld<CPT>(a, o);
              tmem_wait_ld();
#pragma unroll
              for (int k = 0; k < CPT; ++k) o[k] *= sm.fac[cb + k];
              tmem_st<CPT>(a, o);
            }
            tmem_wait_st();
#pragma unroll
            for (int q = 0; q < CPT; ++q) lt[q] *= sm.fac[cb + q];
          }
        }
        // ---- P = exp2(z - m_ref): row sums from the f32 values, bf16 P to shared memory (K-major SW128) ----
        uint8_t* pchunk = s_p + (tl >> 6) * (HB * 128) + (tl & 7) * 2;
        const int tc = (tl & 63) >> 3;
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const float u = ex2_approx(x[q] - sm.mref[cb + q]);
          lt[q] += u;
          const int row = cb + q;
          *reinterpret_cast<__nv_bfloat16*>(pchunk + (row >> 3) * 1024 + (row & 7) * 128 + ((tc ^ (row & 7)) << 4)) =
              __float2bfloat16_rn(u);
        }
        // rows past the end of the sequence feed P.V with P = 0; make sure V there is finite (0 * NaN).
        // Reference: fill_oob_V (the FP8 fork's splitkv_mla.cu:698-721 has the same step).
        if (!tok_valid) {
          const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 15) / 16;
          if (tl < nk * 16) {
            for (int c = cs; c < 8; c += kCS) {
              uint4* row = reinterpret_cast<uint4*>(s_kv + c * kChunkBytes + tl * 128);
#pragma unroll
              for (int k = 0; k < 8; ++k) row[k] = make_uint4(0, 0, 0, 0);
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&sm.p_full);
      }

      // ---- epilogue of this request: O^T / L -> out (or split partials) ----
#pragma unroll
      for (int q = 0; q < CPT; ++q) {
        float v = lt[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sm.red[quad][cb + q] = v;
      }
      named_bar_sync(1, kSmThreads);
      if (st_id < HB) sm.lsum[st_id] = (sm.red[0][st_id] + sm.red[1][st_id]) + (sm.red[2][st_id] + sm.red[3][st_id]);
      named_bar_sync(1, kSmThreads);
      if (sg.n_tiles > 0) {
        mbar_wait(&sm.o_ready, o_seg & 1);
        tc_fence_after();
      }
      const int64_t out_row0 = (int64_t)b * p.M + q_row0;
      if (st_id < HB) {
        const float l = sm.lsum[st_id];
        sm.fac[st_id] = (l == 0.f || l != l) ? 1.f : 1.f / l;
      }
      named_bar_sync(1, kSmThreads);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float o[CPT];
        if (sg.n_tiles > 0) {
          tmem_ld<CPT>(tmem + kColO + j * HB + cb + lane_addr, o);
          tmem_wait_ld();
        } else {
#pragma unroll
          for (int k = 0; k < CPT; ++k) o[k] = 0.f;
        }
        const int d = j * 128 + tl;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int q = cb + k;
          if (q < n_valid_q) {
            const float val = o[k] * sm.fac[q];
            if (sg.no_split) p.out[(out_row0 + q) * 512 + d] = __float2bfloat16_rn(val);
            else p.o_accum[((int64_t)sg.split_idx * p.M + q_row0 + q) * 512 + d] = val;
          }
        }
      }
      if (st_id < n_valid_q) {
        const int row = q_row0 + st_id;
        const float l = sm.lsum[st_id];
        const bool bad = (l == 0.f || l != l);
        const float lse2 = log2f(l) + sm.mref[st_id];
        if (sg.no_split) {
          const int s = row / p.h_q, h = row % p.h_q;
          p.lse[((int64_t)b * p.h_q + h) * p.s_q + s] = bad ? INFINITY : lse2 * 0.6931471805599453f;
        } else {
          p.lse_accum[(int64_t)sg.split_idx * p.M + row] = bad ? -INFINITY : lse2;
        }
      }
      if (sg.n_tiles > 0) {
        tc_fence_before();
        mbar_arrive(&sm.o_free);
        ++o_seg;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace bf16kv
}  // namespace b200

using namespace b200;

extern "C" int b200_mla_decode_bf16(const void* q, const void* k_cache, int64_t num_pages, const int32_t* block_table,
                                    int64_t block_table_stride, const int32_t* cache_seqlens,
                                    const int32_t* tile_scheduler_metadata, const int32_t* num_splits,
                                    int num_sm_parts, int batch, int s_q, int h_q, float softmax_scale, int is_causal,
                                    void* out, float* lse, void* workspace, size_t workspace_bytes, int use_pdl,
                                    void* stream) {
  if (batch == 0) return 0;  // IDLE batch
  B200_CHECK(q && k_cache && block_table && cache_seqlens && tile_scheduler_metadata && num_splits && out && lse &&
                 workspace,
             "mla_decode_bf16: null pointer argument");
  B200_CHECK(batch > 0 && s_q > 0 && h_q > 0 && num_pages > 0, "mla_decode_bf16: batch=%d s_q=%d h_q=%d num_pages=%lld",
             batch, s_q, h_q, (long long)num_pages);
  B200_CHECK(num_sm_parts > 0, "mla_decode_bf16: num_sm_parts must be positive");
  B200_CHECK(workspace_bytes >= b200_mla_workspace_bytes(batch, s_q, h_q, num_sm_parts),
             "mla_decode_bf16: workspace too small (%zu < %zu)", workspace_bytes,
             b200_mla_workspace_bytes(batch, s_q, h_q, num_sm_parts));
  B200_CHECK((reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k_cache) & 15) == 0,
             "mla_decode_bf16: q / k_cache must be 16-byte aligned");
  if (s_q == 1) is_causal = 0;  // flash_api.cpp:96
  const int M = s_q * h_q;
  const int q_blocks = (M + bf16kv::HB - 1) / bf16kv::HB;

  MlaParams prm;
  memset(&prm, 0, sizeof(prm));
  const uint64_t q_rows = (uint64_t)batch * M;
  const uint64_t tokens = (uint64_t)num_pages * kPage;
  if (make_tma_2d(&prm.tm_q_nope, q, 2, 576, q_rows, 1152, 64, bf16kv::HB, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_k_nope, k_cache, 2, 576, tokens, 1152, 64, kPage, TmaSwizzle::B128)) return 1;
  prm.block_table = block_table;
  prm.block_table_stride = block_table_stride;
  prm.seqlens = cache_seqlens;
  prm.meta = tile_scheduler_metadata;
  prm.num_splits = num_splits;
  prm.out = (__nv_bfloat16*)out;
  prm.lse = lse;
  const size_t splits = (size_t)batch + num_sm_parts;
  prm.o_accum = (float*)workspace;
  prm.lse_accum = (float*)workspace + splits * M * 512;
  prm.batch = batch;
  prm.s_q = s_q;
  prm.h_q = h_q;
  prm.M = M;
  prm.scale_log2 = softmax_scale * 1.4426950408889634f;
  prm.is_causal = is_causal;

  const int smem = bf16kv::kSmemTotal + 1024;
  // per device and context, so not cached per process (a process may drive several GPUs)
  B200_CUDA(cudaFuncSetAttribute(bf16kv::mla_decode_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  if (int rc = launch_kernel(bf16kv::mla_decode_bf16_kernel, dim3(q_blocks, 1, num_sm_parts), dim3(bf16kv::kThreads), smem,
                             (cudaStream_t)stream, use_pdl != 0, prm))
    return rc;
  return b200_mla_combine(prm.o_accum, prm.lse_accum, num_splits, out, lse, batch, s_q, h_q, use_pdl, stream);
}
