// Paged MLA decode attention over the per-token-FP8 KV pool, sm_100a (tcgen05 + TMEM + TMA).
//
// Replaces ckv_fp8_per_token::flash_fwd_splitkv_mla_kernel of the reference
// (3rdparty/flashmla-fp8/csrc/kernels/ckv_fp8_per_token/splitkv_mla.cu:1147-1515, host side
// csrc/flash_api.cpp:216-397).  Same split-KV work partition (tile-scheduler metadata), same arithmetic contract
// (e4m3 x e4m3 -> f32 for the 512 "nope" dims + bf16 x bf16 -> same f32 accumulator for the 64 rope dims, logits
// scaled by q_scale * k_scale, V-dequant scale folded into P, P re-quantised to e4m3 before P.V), different
// machine mapping:
//
//   SwapAB everywhere: KV tokens sit on the UMMA M side (128 tokens = 2 pages per tile), the q rows of one request
//   (HB = 16/32/64 per CTA) on the UMMA N side.
//       S^T[128 tok, HB]  = K_nope[128, 512] . Qn^T  (16 x kind::f8f6f4, K-major SW128, straight from TMA)
//                         + K_rope[128, 64]  . Qr^T  ( 4 x kind::f16 into the same TMEM accumulator)
//       O^T[512 d, HB]   += V^T[512, 128 tok] . P[128 tok, HB]
//   V^T is the *same* shared-memory bytes as K_nope, described MN-major to the tensor core -- the reference's
//   LDSM/byte_perm/STSM transpose of V (ckv_fp8_per_token/utils.h:39-147) disappears.
//   The four 128-byte k-chunks of the QK contraction are exactly the four 128-lane M tiles of the PV product.
//
//   Warp roles (1 CTA/SM): warp 0 = TMA producer, warp 1 = UMMA issuer for Q.K^T, warp 2 = UMMA issuer for P.V,
//   warp 3 idle, warps 4.. = softmax / correction / epilogue (4 TMEM lane quadrants x HB/CPT column slices; one
//   thread per (KV token, column slice)).
//   TMEM: S double-buffered (2 x HB columns) + O (4 d-chunks x [hi | lo] = 8 x HB columns).
//
//   Softmax: the reference's page-wise FP8 codes of P under a lazily moved per-row reference (see the comment above
//   mla_decode_kernel); the row sum L is kept per token-lane in registers and reduced once per request.
#include <stdlib.h>
#include <string.h>

#include "mla_common.cuh"

namespace b200 {

static constexpr int kStages = 2;
static constexpr int kNopeChunkBytes = kTileTokens * 128;     // one 128-byte k-chunk of a tile: 16 KB
static constexpr int kStageNopeBytes = 4 * kNopeChunkBytes;   // 64 KB
static constexpr int kStageRopeBytes = kTileTokens * 128;     // 16 KB
static constexpr int kStageBytes = kStageNopeBytes + kStageRopeBytes;
static constexpr int kPageTxBytes = kPage * 512 + kPage * 128;

template <int HB>
struct SmemLayout {
  static constexpr int kQBytes = HB * 640;           // 4 nope chunks [HB x 128 B] + rope [HB x 128 B]
  static constexpr int kPBytes = HB * kTileTokens;   // [HB/16 chunks][16 groups][8 tokens][16 q]
  // P buffers: (hi, lo) pairs by tile parity; the lo tile sits right behind the hi tile, so the pair is ONE B operand
  // of 2 * HB columns
  static constexpr int kPPairs = 2;
  static constexpr int kOffKV = 0;
  static constexpr int kOffQ = kStages * kStageBytes;
  static constexpr int kOffP = kOffQ + kQBytes;
  static constexpr int kOffSmall = kOffP + kPPairs * 2 * kPBytes;
  static constexpr int kTotal = kOffSmall + 4096;
};

struct SmallSmem {
  uint64_t kv_full[kStages], kv_empty[kStages];
  uint64_t q_full, q_empty;
  uint64_t s_full[2], p_full[2];
  uint64_t pv_done, o_ready, o_free;
  uint32_t tmem_base;
  float f[64];          // q_scale * softmax_scale * log2(e)
  float rfin[64];       // per q row: final reference exponent of the request (epilogue)
  float inv[64];        // per q row: 1 / L
  float lsum[64];
  int32_t lim[64];      // per q row: first masked token index
  float red[2][4][64];  // [tile parity][TMEM lane quadrant][q row]: per-warp column maxima / row-sum partials
};
static_assert(sizeof(SmallSmem) <= 4096, "SmallSmem too large");

// HB  = q rows per CTA (UMMA N); CPT = S columns (q rows) each softmax thread owns.
// The softmax group has 4 * HB/CPT warps: 4 TMEM lane quadrants x HB/CPT column slices; one thread per (KV token,
// column slice).
//
// Softmax arithmetic (per q row, per 64-token page -- the reference's granularity, splitkv_mla.cu:391-458):
//   y_j     = s_j * q_scale * k_scale_j * scale_log2 + log2(k_scale_j)       (log2 of the un-normalised P'_j)
//   ymax    = max_j y_j over the page                    -> the reference's p_scale * 2^m = 2^ymax / 448
//   code_j  = e4m3(448 * 2^(y_j - ymax))                 -> the reference's P8 codes (:433-440), the page's top = 448
//   sum     = sum_j 448 * 2^(y_j - ymax) / k_scale_j     -> the reference's cur_sum / p_scale (:426,434)
// The reference then rescales O and L by exp2(m_old - m_new) * p_old / p_new every page (:442-455); with O in TMEM that
// is a 128 KB read-modify-write per page.  Here O and L live under a per-row reference exponent R that only moves when
// a page's ymax exceeds it (then R = ymax + kRefHeadroom and O is rescaled once), and the page's weight w = 2^(ymax - R) <= 1
// multiplies the codes instead: code_j * w as an e4m3 pair hi + lo, 8 significant bits -- the result differs from the
// page-wise recurrence by ~1e-3 (rounding of w*code, fp32 summation order), not by the 2^-4 noise of a different e4m3
// grid.  Under SwapAB P is the B operand and HB is small, so hi and lo ride in ONE product: B = [P_hi | P_lo] with
// N = 2 HB columns, O^T = [O_hi | O_lo] in TMEM, added in the epilogue -- P.V issues as many UMMAs (and reads V^T from
// shared memory as often) as with a single P tile.
template <int HB, int CPT>
__global__ void __launch_bounds__(128 + 128 * (HB / CPT), 1) mla_decode_kernel(const __grid_constant__ MlaParams p) {
  using L = SmemLayout<HB>;
  static_assert(HB == 16 || HB == 32, "HB must be 16 or 32 (larger M runs on the CTA-pair kernel)");
  constexpr uint32_t kTmemCols = HB == 32 ? 512 : 256;
  constexpr uint32_t kColS = 0, kColO = 2 * HB;  // S: 2 x HB columns; O: 4 d-chunks x [hi HB | lo HB] columns
  constexpr int kCS = HB / CPT;                 // column slices
  constexpr int kSmThreads = 128 * kCS;         // softmax group size
  static_assert(CPT % 8 == 0 && CPT <= 32, "CPT must be 8, 16 or 32");

  extern __shared__ uint8_t smem_raw[];
  // round up to 1024 B by pointer arithmetic (an integer round trip would demote every access to generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_kv = smem + L::kOffKV;
  uint8_t* s_q = smem + L::kOffQ;
  uint8_t* s_p = smem + L::kOffP;
  SmallSmem& sm = *reinterpret_cast<SmallSmem*>(smem + L::kOffSmall);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q_block = blockIdx.x, part = blockIdx.z;
  const int q_row0 = q_block * HB;
  const int n_valid_q = min(HB, p.M - q_row0);

  const int32_t* meta = p.meta + part * B200_MLA_METADATA_INTS;
  const int begin_idx = __ldg(meta + 0), begin_seqlen = __ldg(meta + 1), end_idx = __ldg(meta + 2),
            end_seqlen = __ldg(meta + 3), begin_n_split = __ldg(meta + 4);
  // the split-KV combine kernel behind this one is launched with programmatic stream serialization: its CTAs may be
  // scheduled as soon as SMs drain (its griddepcontrol.wait orders the reads of the partials)
  pdl_launch_dependents();
  if (begin_idx >= p.batch) return;  // splitkv_mla.cu:1226

  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&sm.kv_full[i], 1);
      mbar_init(&sm.kv_empty[i], 1);
    }
    mbar_init(&sm.q_full, 1);
    mbar_init(&sm.q_empty, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.p_full[i], kSmThreads);
    }
    mbar_init(&sm.pv_done, 1);
    mbar_init(&sm.o_ready, 1);
    mbar_init(&sm.o_free, kSmThreads);
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_q_nope);
    tma_prefetch_desc(&p.tm_q_rope);
    tma_prefetch_desc(&p.tm_k_nope);
    tma_prefetch_desc(&p.tm_k_rope);
  }
  if (warp == 1) tmem_alloc(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  // Issuer warps run warp-uniform code (every lane executes the loops and the barrier waits) and predicate only the
  // TMA / MMA instructions with elect.sync: descriptors and coordinates then live in uniform registers.  Issuing from
  // a divergent `if (lane == 0)` costs ~110 cycles per tcgen05.mma (R2UR traffic), measured with b200_umma_rate.
  if (warp == 0) {
    // ============================== TMA producer ==============================
    const uint64_t pol_kv = policy_evict_first();
    const bool shared_kv = gridDim.x > 1;  // sibling q-block CTAs stream the same pages: let L2 keep them
    uint32_t g = 0, seg_i = 0;
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      if (sg.n_tiles == 0) continue;
      // Q tile of this request (rows beyond the request are masked by the consumers)
      mbar_wait(&sm.q_empty, (seg_i & 1) ^ 1);
      ++seg_i;
      const int qrow = b * p.M + q_row0;
      if (elect_one()) {
        mbar_arrive_expect_tx(&sm.q_full, L::kQBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d(s_q + c * HB * 128, &p.tm_q_nope, &sm.q_full, c * 128, qrow);
        tma_load_2d(s_q + 4 * HB * 128, &p.tm_q_rope, &sm.q_full, 0, qrow);
      }
      __syncwarp();
      const int32_t* bt = p.block_table + (int64_t)b * p.block_table_stride;
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const uint32_t st = g % kStages;
        mbar_wait(&sm.kv_empty[st], ((g / kStages) & 1) ^ 1);
        const int blk0 = sg.start_block + 2 * t;
        const int npages = min(2, sg.end_block - blk0);
        uint8_t* stage = s_kv + st * kStageBytes;
        const int tok0 = __ldg(bt + blk0) * kPage;
        const int tok1 = npages > 1 ? __ldg(bt + blk0 + 1) * kPage : 0;
        // warm L2 for the tile that will reuse this stage (two tiles ahead): HBM latency is then overlapped with
        // the consumption of the current tiles instead of being paid after the stage is released
        const int pf0 = blk0 + 2 * kStages < sg.end_block ? __ldg(bt + blk0 + 2 * kStages) * kPage : -1;
        const int pf1 = blk0 + 2 * kStages + 1 < sg.end_block ? __ldg(bt + blk0 + 2 * kStages + 1) * kPage : -1;
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.kv_full[st], npages * kPageTxBytes);
          if (shared_kv) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              tma_load_2d(stage + c * kNopeChunkBytes, &p.tm_k_nope, &sm.kv_full[st], c * 128, tok0);
            tma_load_2d(stage + kStageNopeBytes, &p.tm_k_rope, &sm.kv_full[st], 0, tok0);
            if (npages > 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                tma_load_2d(stage + c * kNopeChunkBytes + kPage * 128, &p.tm_k_nope, &sm.kv_full[st], c * 128, tok1);
              tma_load_2d(stage + kStageNopeBytes + kPage * 128, &p.tm_k_rope, &sm.kv_full[st], 0, tok1);
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c)
              tma_load_2d_hint(stage + c * kNopeChunkBytes, &p.tm_k_nope, &sm.kv_full[st], c * 128, tok0, pol_kv);
            tma_load_2d_hint(stage + kStageNopeBytes, &p.tm_k_rope, &sm.kv_full[st], 0, tok0, pol_kv);
            if (npages > 1) {
#pragma unroll
              for (int c = 0; c < 4; ++c)
                tma_load_2d_hint(stage + c * kNopeChunkBytes + kPage * 128, &p.tm_k_nope, &sm.kv_full[st], c * 128,
                                 tok1, pol_kv);
              tma_load_2d_hint(stage + kStageNopeBytes + kPage * 128, &p.tm_k_rope, &sm.kv_full[st], 0, tok1, pol_kv);
            }
          }
          if (pf0 >= 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_prefetch_2d(&p.tm_k_nope, c * 128, pf0);
            tma_prefetch_2d(&p.tm_k_rope, 0, pf0);
          }
          if (pf1 >= 0) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_prefetch_2d(&p.tm_k_nope, c * 128, pf1);
            tma_prefetch_2d(&p.tm_k_rope, 0, pf1);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ============================== UMMA issuer: S^T = K . Q^T ==============================
    constexpr uint32_t idesc_qk8 = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, HB);
    constexpr uint32_t idesc_qk16 = make_idesc(FMT_BF16, FMT_BF16, 0, 0, 128, HB);
    const uint32_t kv_addr = smem_u32(s_kv), q_addr = smem_u32(s_q);
    const uint64_t dk_sw128 = make_smem_desc(0, 0, 1024, SWZ_128B);  // K-major, 128-B swizzle
    auto with_addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3FFF); };
    uint32_t g = 0, seg = 0;
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      if (sg.n_tiles == 0) continue;
      mbar_wait(&sm.q_full, seg & 1);
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const uint32_t st = g % kStages, buf = g & 1;
        mbar_wait(&sm.kv_full[st], (g / kStages) & 1);
        mbar_wait(&sm.p_full[buf], ((g >> 1) & 1) ^ 1);  // S[buf] drained by the softmax of tile g-2
        tc_fence_after();
        const uint32_t stage = kv_addr + st * kStageBytes;
        const uint32_t d = tmem + kColS + buf * HB;
        if (elect_one()) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_f8(d, with_addr(dk_sw128, stage + c * kNopeChunkBytes + s * 32),
                      with_addr(dk_sw128, q_addr + c * HB * 128 + s * 32), idesc_qk8, (c | s) != 0);
#pragma unroll
          for (int s = 0; s < 4; ++s)
            umma_f16(d, with_addr(dk_sw128, stage + kStageNopeBytes + s * 32),
                     with_addr(dk_sw128, q_addr + 4 * HB * 128 + s * 32), idesc_qk16, 1);
          umma_commit(&sm.s_full[buf]);
          // last Q.K^T of this request issued: Q smem is reusable once these MMAs retire
          if (t == sg.n_tiles - 1) umma_commit(&sm.q_empty);
        }
        __syncwarp();
      }
      ++seg;
    }
  } else if (warp == 2) {
    // ============================== UMMA issuer: O^T += V^T . (P_hi + P_lo) ==============================
    constexpr uint32_t idesc_pv = make_idesc(FMT_E4M3, FMT_E4M3, 1, 1, 128, 2 * HB);  // N = [P_hi | P_lo]
    const uint32_t kv_addr = smem_u32(s_kv), p_addr = smem_u32(s_p);
    const uint64_t dv_sw128 = make_smem_desc(0, 16384, 1024, SWZ_128B);  // MN-major view of the K tile bytes
    const uint64_t dp_none = make_smem_desc(0, 128, 2048, SWZ_NONE);     // P: 8x16-B core matrices
    auto with_addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3FFF); };
    uint32_t g = 0, seg = 0;
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      if (sg.n_tiles == 0) continue;
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const uint32_t st = g % kStages, buf = g & 1;
        mbar_wait(&sm.p_full[buf], (g >> 1) & 1);
        if (t == 0 && seg > 0) mbar_wait(&sm.o_free, (seg - 1) & 1);
        tc_fence_after();
        const uint32_t stage = kv_addr + st * kStageBytes;
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 31) / 32;
        const uint32_t pbuf = p_addr + buf * (2 * L::kPBytes);
        if (elect_one()) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int s = 0; s < 4; ++s)
              if (s < nk)
                umma_f8(tmem + kColO + j * (2 * HB), with_addr(dv_sw128, stage + j * kNopeChunkBytes + s * 4096),
                        with_addr(dp_none, pbuf + s * 512), idesc_pv, (t | s) != 0);
          umma_commit(&sm.kv_empty[st]);
          umma_commit(&sm.pv_done);
          if (t == sg.n_tiles - 1) umma_commit(&sm.o_ready);  // one completion per request: the epilogue's gate
        }
        __syncwarp();
      }
      ++seg;
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue ==============================
    const int st_id = tid - 128;                // 0 .. kSmThreads-1 within the softmax group
    const int quad = warp & 3;                  // TMEM lane quadrant this warp may access
    const int cs = (warp - 4) >> 2;             // column slice
    const int cb = cs * CPT;                    // first q column of this thread
    const int tl = quad * 32 + lane;            // token lane within the tile
    const int page = tl >> 6;                   // which of the tile's two pages this token lane belongs to
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    uint32_t g = 0, o_seg = 0;
    // per q column of this thread (identical in every thread that owns the column: they are functions of the shared
    // maxima only): R = reference exponent, lt = this token lane's partial row sum in units of 2^R / 448
    float R[CPT], lt[CPT];

    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment(p, b, begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, HB);
      // ---- per-request state ----
      named_bar_sync(1, kSmThreads);  // previous request's readers of sm.* are done
      if (st_id < HB) {
        const int row = q_row0 + st_id;
        const bool vq = st_id < n_valid_q;
        sm.f[st_id] = vq ? __ldg(p.q_scale + (int64_t)b * p.M + row) * p.scale_log2 : 0.f;
        int lim = 0;
        if (vq) lim = p.is_causal ? min(sg.seqlen - mask_len(p.s_q, p.h_q, p.M, row), sg.end_block * kPage) : sg.valid_end;
        sm.lim[st_id] = lim;
      }
      named_bar_sync(1, kSmThreads);
#pragma unroll
      for (int q = 0; q < CPT; ++q) R[q] = kMInit, lt[q] = 0.f;
      const int32_t* bt = p.block_table + (int64_t)b * p.block_table_stride;
      const bool all_q_valid = n_valid_q == HB && !p.is_causal;

      // this token lane's k scale, fetched one tile ahead (two dependent L2 / HBM round trips otherwise sit between
      // s_full and the first use)
      auto fetch_ks = [&](int t) -> float {
        if (t >= sg.n_tiles) return 1.f;
        const int token = (sg.start_block + 2 * t) * kPage + tl;
        if (token >= sg.valid_end) return 1.f;
        const int pg = __ldg(bt + sg.start_block + 2 * t + page);
        return __ldg(p.k_scale + (int64_t)pg * kPage + (tl & 63));
      };
      float ks_next = fetch_ks(0);

      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        const uint32_t buf = g & 1;
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        const int token = tile_start + tl;
        const bool tok_valid = token < sg.valid_end;
        const float ks = ks_next;
        ks_next = fetch_ks(t + 1);
        const float lk = log2f(ks), rk = 1.f / ks;
        const bool full_tile = all_q_valid && (tile_start + kTileTokens <= sg.valid_end);

        mbar_wait(&sm.s_full[buf], (g >> 1) & 1);
        tc_fence_after();
        const uint32_t s_addr = tmem + kColS + buf * HB + cb + lane_addr;
        // y = log2 of the un-normalised P' of this token for 8 q columns (S is read from TMEM in both passes: nothing
        // but R and lt stays live across the barrier)
        auto load_y8 = [&](int c8, float* y) {
          tmem_ld<8>(s_addr + c8 * 8, y);
          tmem_wait_ld();
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = cb + c8 * 8 + q;
            const float v = fmaf(y[q] * sm.f[col], ks, lk);
            const bool ok = tok_valid && (full_tile || token < sm.lim[col]);
            y[q] = ok ? v : -INFINITY;  // a select, not arithmetic: S of a poisoned (NaN) padding row never propagates
          }
        };

        // ---- pass A: page maxima per q column ----
        float (*red)[64] = sm.red[g & 1];
#pragma unroll
        for (int c8 = 0; c8 < CPT / 8; ++c8) {
          float y[8];
          load_y8(c8, y);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int col = cb + c8 * 8 + q;
            const float wm = warp_max_f32(y[q]);
            if ((col & 31) == lane) red[quad][col] = wm;
          }
        }
        named_bar_sync(1, kSmThreads);

        // ---- reference moves (per q column; every thread that owns the column takes the same decision) ----
        bool any_move = false;
#pragma unroll
        for (int q = 0; q < CPT; ++q) {
          const float yb = fmaxf(fmaxf(red[0][cb + q], red[1][cb + q]), fmaxf(red[2][cb + q], red[3][cb + q]));
          any_move |= yb > R[q];
        }
        if (any_move) {  // warp-uniform
          if (t != 0) {
            // every P.V up to the previous tile must have landed in O before it is rescaled
            mbar_wait(&sm.pv_done, (g - 1) & 1);
            tc_fence_after();
          }
#pragma unroll
          for (int c8 = 0; c8 < CPT / 8; ++c8) {
            float fac[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int col = cb + c8 * 8 + q;
              const float yb = fmaxf(fmaxf(red[0][col], red[1][col]), fmaxf(red[2][col], red[3][col]));
              const float rn = yb > R[c8 * 8 + q] ? yb + kRefHeadroom : R[c8 * 8 + q];
              fac[q] = exp2f(R[c8 * 8 + q] - rn);  // 0 on a row's first move (R = -1e30), 1 for rows that stay
              R[c8 * 8 + q] = rn;
              lt[c8 * 8 + q] *= fac[q];
            }
            if (t != 0) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {  // 4 d-chunks x (hi, lo)
                float o[8];
                const uint32_t a = tmem + kColO + j * HB + cb + c8 * 8 + lane_addr;
                tmem_ld<8>(a, o);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] *= fac[k];
                tmem_st<8>(a, o);
              }
            }
          }
          if (t != 0) tmem_wait_st();
        }

        // ---- pass B: page codes, page weight, hi / lo split ----
        uint32_t hi[CPT / 4], lo[CPT / 4];
#pragma unroll
        for (int c8 = 0; c8 < CPT / 8; ++c8) {
          float y[8];
          load_y8(c8, y);
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            float w[2], v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int col = cb + c8 * 8 + q + e;
              const float m = page ? fmaxf(red[2][col], red[3][col]) : fmaxf(red[0][col], red[1][col]);
              const bool live = m > -INFINITY;
              w[e] = live ? ex2_approx(m - R[c8 * 8 + q + e]) : 0.f;
              v[e] = ex2_approx(y[q + e] - (live ? m : 0.f) + kLog2Fp8Max);
            }
            const uint32_t w_h2 = f32x2_to_f16x2(w[0], w[1]);
            lt[c8 * 8 + q] = fmaf(f16_lo_to_f32(w_h2), v[0] * rk, lt[c8 * 8 + q]);
            lt[c8 * 8 + q + 1] = fmaf(f16_hi_to_f32(w_h2), v[1] * rk, lt[c8 * 8 + q + 1]);
            uint32_t h2, l2;
            split_hi_lo(cvt_e4m3x2(v[0], v[1]), w_h2, h2, l2);
            if ((q & 2) == 0) hi[c8 * 2 + q / 4] = h2, lo[c8 * 2 + q / 4] = l2;
            else hi[c8 * 2 + q / 4] |= h2 << 16, lo[c8 * 2 + q / 4] |= l2 << 16;
          }
        }

        // ---- publish P (B operand of P.V: [token][q], MN-major, 8x16-byte core matrices) ----
        // The P pair (and this KV stage) was last read by P.V of tile g-2 (two pairs) / g-1 (one pair), issued by
        // another warp: wait for it to retire
        mbar_wait(&sm.kv_empty[g % kStages], ((g / kStages) & 1) ^ 1);
        uint8_t* prow = s_p + buf * (2 * L::kPBytes) + (cb >> 4) * 2048 + (tl >> 3) * 128 + (tl & 7) * 16 + (cb & 15);
        if constexpr (CPT == 8) {
          *reinterpret_cast<uint2*>(prow) = make_uint2(hi[0], hi[1]);
          *reinterpret_cast<uint2*>(prow + L::kPBytes) = make_uint2(lo[0], lo[1]);
        } else {
#pragma unroll
          for (int c = 0; c < CPT / 16; ++c) {
            *reinterpret_cast<uint4*>(prow + c * 2048) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
            *reinterpret_cast<uint4*>(prow + c * 2048 + L::kPBytes) =
                make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
          }
        }
        // rows past the end of the sequence feed P.V with P = 0; make sure V there is finite (0 * NaN).
        // Reference: fill_oob_V, splitkv_mla.cu:698-721.
        if (!tok_valid && cs == 0) {
          const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 31) / 32;
          if (tl < nk * 32) {
            uint8_t* stage = s_kv + (g % kStages) * kStageBytes;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              uint4* row = reinterpret_cast<uint4*>(stage + c * kNopeChunkBytes + tl * 128);
#pragma unroll
              for (int k = 0; k < 8; ++k) row[k] = make_uint4(0, 0, 0, 0);
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        mbar_arrive(&sm.p_full[buf]);
      }

      // ---- epilogue of this request: O^T / L -> out (or split partials) ----
      float (*red)[64] = sm.red[0];
      named_bar_sync(1, kSmThreads);  // the last tile's readers of red[0] are done
#pragma unroll
      for (int q = 0; q < CPT; ++q) {
        float v = lt[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[quad][cb + q] = v;
        if (tl == 0) sm.rfin[cb + q] = R[q];
      }
      named_bar_sync(1, kSmThreads);
      // fixed summation order: results are bit-reproducible run to run
      if (st_id < HB) {
        const float l = (red[0][st_id] + red[1][st_id]) + (red[2][st_id] + red[3][st_id]);
        sm.lsum[st_id] = l;
        sm.inv[st_id] = (l == 0.f || l != l) ? 1.f : 1.f / l;  // splitkv_mla.cu:1464-1466
      }
      named_bar_sync(1, kSmThreads);
      if (sg.n_tiles > 0) {
        // not pv_done: it completes once per tile and a parity wait from here could be two phases behind
        mbar_wait(&sm.o_ready, o_seg & 1);
        tc_fence_after();
      }
      const int64_t out_row0 = (int64_t)b * p.M + q_row0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float o[CPT];
        if (sg.n_tiles > 0) {
          float ol[CPT];
          tmem_ld<CPT>(tmem + kColO + j * (2 * HB) + cb + lane_addr, o);
          tmem_ld<CPT>(tmem + kColO + j * (2 * HB) + HB + cb + lane_addr, ol);
          tmem_wait_ld();
#pragma unroll
          for (int k = 0; k < CPT; ++k) o[k] += ol[k];  // O = O_hi + O_lo
        } else {
#pragma unroll
          for (int k = 0; k < CPT; ++k) o[k] = 0.f;
        }
        const int d = j * 128 + tl;
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int q = cb + k;
          if (q < n_valid_q) {
            const float val = o[k] * sm.inv[q];
            if (sg.no_split) p.out[(out_row0 + q) * 512 + d] = __float2bfloat16_rn(val);
            else p.o_accum[((int64_t)sg.split_idx * p.M + q_row0 + q) * 512 + d] = val;
          }
        }
      }
      if (st_id < n_valid_q) {
        const int row = q_row0 + st_id;
        const float l = sm.lsum[st_id];
        const bool bad = (l == 0.f || l != l);
        const float lse2 = log2f(l) + (sm.rfin[st_id] - kLog2Fp8Max);  // L_true = L * 2^R / 448
        if (sg.no_split) {
          const int s = row / p.h_q, h = row % p.h_q;
          p.lse[((int64_t)b * p.h_q + h) * p.s_q + s] = bad ? INFINITY : lse2 * 0.6931471805599453f;  // :1484
        } else {
          p.lse_accum[(int64_t)sg.split_idx * p.M + row] = bad ? -INFINITY : lse2;                    // :1506
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

template <int HB, int CPT>
static int launch_mla(const MlaParams& prm, int q_blocks, int num_sm_parts, cudaStream_t stream) {
  const int smem = SmemLayout<HB>::kTotal + 1024;
  B200_CUDA(cudaFuncSetAttribute(mla_decode_kernel<HB, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  mla_decode_kernel<HB, CPT><<<dim3(q_blocks, 1, num_sm_parts), 128 + 128 * (HB / CPT), smem, stream>>>(prm);
  B200_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" {

size_t b200_mla_merge_state_ints(int batch, int s_q, int h_q) {
  // {ticket, done} per (request, 128-row q block, CTA of the pair)
  if (batch <= 0 || s_q <= 0 || h_q <= 0) return 0;
  return (size_t)batch * ((s_q * h_q + 127) / 128) * 4;
}

size_t b200_mla_workspace_bytes(int batch, int s_q, int h_q, int num_sm_parts) {
  const size_t M = (size_t)s_q * h_q;
  const size_t splits = (size_t)batch + num_sm_parts;
  return splits * M * 512 * sizeof(float) + splits * M * sizeof(float);
}

int b200_mla_decode_ckv_fp8_per_token(const void* q_nope, const void* q_rope, const float* q_scale, const void* k_lora,
                                      const void* k_rope, const float* k_scale, int64_t num_pages,
                                      const int32_t* block_table, int64_t block_table_stride,
                                      const int32_t* cache_seqlens, const int32_t* tile_scheduler_metadata,
                                      const int32_t* num_splits, int num_sm_parts, int batch, int s_q, int h_q,
                                      float softmax_scale, int is_causal, void* out, float* lse, void* workspace,
                                      size_t workspace_bytes, int32_t* merge_state, int use_pdl, void* stream) {
  if (batch == 0) return 0;  // IDLE batch (flash_api.cpp accepts batch_size 0): nothing to attend
  B200_CHECK(q_nope && q_rope && q_scale && k_lora && k_rope && k_scale && block_table && cache_seqlens &&
                 tile_scheduler_metadata && num_splits && out && lse && workspace,
             "mla_decode: null pointer argument");
  B200_CHECK(batch > 0 && s_q > 0 && h_q > 0 && num_pages > 0, "mla_decode: batch=%d s_q=%d h_q=%d num_pages=%lld", batch,
             s_q, h_q, (long long)num_pages);
  const int M = s_q * h_q;
  const int q_blocks = (M + B200_MLA_BLOCK_M - 1) / B200_MLA_BLOCK_M;
  B200_CHECK(num_sm_parts > 0, "mla_decode: num_sm_parts must be positive");
  B200_CHECK(workspace_bytes >= b200_mla_workspace_bytes(batch, s_q, h_q, num_sm_parts),
             "mla_decode: workspace too small (%zu < %zu)", workspace_bytes,
             b200_mla_workspace_bytes(batch, s_q, h_q, num_sm_parts));
  B200_CHECK((reinterpret_cast<uintptr_t>(k_lora) & 15) == 0 && (reinterpret_cast<uintptr_t>(k_rope) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(q_nope) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_rope) & 15) == 0,
             "mla_decode: q / k tensors must be 16-byte aligned");
  if (s_q == 1) is_causal = 0;  // flash_api.cpp:279

  const int HB = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  MlaParams prm;
  memset(&prm, 0, sizeof(prm));
  const uint64_t q_rows = (uint64_t)batch * M;
  const uint64_t tokens = (uint64_t)num_pages * kPage;
  if (make_tma_2d(&prm.tm_q_nope, q_nope, 1, 512, q_rows, 512, 128, HB, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_q_rope, q_rope, 2, 64, q_rows, 128, 64, HB, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_k_nope, k_lora, 1, 512, tokens, 512, 128, kPage, TmaSwizzle::B128)) return 1;
  if (make_tma_2d(&prm.tm_k_rope, k_rope, 2, 64, tokens, 128, 64, kPage, TmaSwizzle::B128)) return 1;
  prm.q_scale = q_scale;
  prm.k_scale = k_scale;
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
  prm.scale_log2 = softmax_scale * 1.4426950408889634f;  // flash_api.cpp:324
  prm.is_causal = is_causal;
  static const int trace = getenv("B200_MLA_TRACE") ? atoi(getenv("B200_MLA_TRACE")) : 0;  // diagnostic bit flags
  prm.trace = trace;
  static const int pf_tiles = getenv("B200_MLA_PF") ? atoi(getenv("B200_MLA_PF")) : 4;  // tuning knob, see DESIGN.md
  prm.pf_tiles = pf_tiles < 2 ? 2 : pf_tiles;
  static const int pf_mode = getenv("B200_MLA_PFMODE") ? atoi(getenv("B200_MLA_PFMODE")) : 2;  // bulk L2 prefetch of the K pages 4 tiles ahead: 2-5 % at the metric shape
  prm.pf_mode = pf_mode;
  prm.k_lora_base = (const uint8_t*)k_lora;
  prm.k_rope_base = (const uint8_t*)k_rope;

  static const int q_pf = getenv("B200_MLA_QPF") ? atoi(getenv("B200_MLA_QPF")) : 6;
  prm.q_pf_tiles = q_pf < 1 ? 1 : q_pf;
  static const int ks_ahead = getenv("B200_MLA_KSAHEAD") ? atoi(getenv("B200_MLA_KSAHEAD")) : 6;
  prm.ks_ahead = ks_ahead < 1 ? 1 : (ks_ahead > 6 ? 6 : ks_ahead);

  int rc;
  // M <= 32: single-CTA SwapAB kernel (KV tokens on UMMA M, q rows on N); above: the CTA-pair kernel (32 < M <= 64 fills
  // one CTA of the pair -- measured faster than a 64-column SwapAB tile)
  if (M > 32) {
    // with a merge-state buffer the pair kernel merges cut requests itself (the part that finishes last): no combine
    static const int no_merge = getenv("B200_MLA_NO_MERGE") ? atoi(getenv("B200_MLA_NO_MERGE")) : 0;  // A/B knob
    // (only when the whole grid is resident at once: with two waves of parts -- 32 < M <= 64 -- the tails of the merging
    // parts cost more than the combine launch, measured)
    const bool one_wave = 2 * ((M + 127) / 128) * num_sm_parts <= device_sm_count();
    prm.merge_state = (no_merge || !one_wave) ? nullptr : reinterpret_cast<int*>(merge_state);
    // at least one request per part: everything is merged in the kernel, no combine launch (mla_common.cuh)
    const bool all_in_kernel = batch >= num_sm_parts;
    prm.merge_limit = all_in_kernel ? 0x7fffffff : kMaxKernelMerge;
    rc = launch_mla_pair(prm, num_sm_parts, use_pdl != 0 && pdl_default(), (cudaStream_t)stream);
    if (rc || (prm.merge_state && all_in_kernel)) return rc;
    if (prm.merge_state)  // what is left for the combine kernel: requests in more than kMaxKernelMerge pieces
      return launch_mla_combine(prm.o_accum, prm.lse_accum, num_splits, out, lse, batch, s_q, h_q, kMaxKernelMerge, use_pdl,
                                stream);
  } else if (HB == 16) rc = launch_mla<16, 8>(prm, q_blocks, num_sm_parts, (cudaStream_t)stream);
  else rc = launch_mla<32, 16>(prm, q_blocks, num_sm_parts, (cudaStream_t)stream);
  if (rc) return rc;
  return b200_mla_combine(prm.o_accum, prm.lse_accum, num_splits, out, lse, batch, s_q, h_q, use_pdl, stream);
}

}  // extern "C"
