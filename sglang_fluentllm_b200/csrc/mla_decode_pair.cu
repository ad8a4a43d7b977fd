// Paged MLA decode attention over the per-token-FP8 KV pool for 64 < M = h_q * s_q: CTA-pair kernel (cluster of two
// CTAs on one TPC, tcgen05 cta_group::2).
//
// Replaces ckv_fp8_per_token::flash_fwd_splitkv_mla_kernel of the reference for the shapes where it runs two 64-row
// q blocks per request (3rdparty/flashmla-fp8/csrc/kernels/ckv_fp8_per_token/splitkv_mla.cu:1147-1515; the two
// warpgroup "seesaw" of docs/20250422-new-kernel-deep-dive.md:27-50).  Same split-KV metadata, same arithmetic contract
// as mla_decode.cu; the machine mapping is the non-swapped one:
//
//   A pair owns 128 q rows of one request (CTA r: rows 64r..64r+63) and walks its KV in 128-token tiles (2 pages).
//       S[128 q, 128 tok]   = Qn[128, 512] . Kn^T   (16 x kind::f8f6f4, M=128 N=128)  + Qr[128, 64] . Kr^T (4 x kind::f16)
//       O[128 q, 512 d]    += P[128 q, 128 tok] . V  ( 2 d-halves x 4 x kind::f8f6f4, M=128 N=256, V MN-major)
//   cta_group::2 splits M across the pair (each CTA keeps its 64 rows of S and O in its own TMEM, as 128 lanes x N/2
//   columns: lanes 0-63 hold the first half of the N columns, lanes 64-127 the second half -- measured with
//   b200_umma_pair_lab) and splits the B operand's N between the two shared memories: for Q.K^T each CTA stages ONE
//   page (64 tokens x 576 dims), for P.V each CTA stages 128 of every 256 output dims for all 128 tokens.  Per SM that
//   is 72 KB of shared-memory fill per 128 tokens x 128 q rows instead of 160 KB for two independent 64-row CTAs, and
//   every MMA reads at most 6 KB of operands per 64 tensor-core cycles, which is what lets the tensor pipe run near
//   its issue rate (single-CTA SS MMAs at N <= 128 are operand-bandwidth bound, see b200_umma_rate).
//
//   Softmax: a thread owns (q row, 64 token columns = one page) -- exactly the reference's (row, page) granularity, so
//   the page maximum, the page's e4m3 codes and the page sum (mla_decode.cu, comment above mla_decode_kernel) are
//   thread-local.  O and L live under a per-row reference exponent R that moves only when a page exceeds it; the page
//   weight 2^(ymax - R) rides on the codes as an e4m3 hi + lo pair (two P tiles, two P.V products per tile).  Three
//   groups of four warps take the tiles round-robin (group = tile index mod 3); a tile's P depends on its predecessor only
//   through R, so the groups meet at a per-tile hand-off (fin) and a version counter; a group that computed under a
//   reference that has since moved recomputes its second pass.  Both passes over S are compact 16-column loops (TMEM
//   loads double-buffered, packed fp32x2 / f16x2 arithmetic, codes written straight into the swizzled P tiles); the
//   first pass writes y back over S in TMEM, so the second one neither recomputes it nor needs the mask.
//
//   Warp roles per CTA (512 threads): warp 0 = TMA producer for Q, K pages (-> leader barrier via cta_group::2 TMA)
//   and the k-scale ring; warp 2 = V producer (it also zeroes the token rows past a request's end before it completes
//   the unit's landing barrier); leader CTA: warp 1 issues every Q.K^T and warp 3 every P.V (two
//   independent instruction streams into the in-order tensor pipe; Q.K^T runs up to four tiles ahead), other CTA:
//   warp 1 relays the local P / O hand-offs to the leader; warp 1 also owns the TMEM allocation; warps 4-15 = three
//   softmax / correction / epilogue groups (group = tile index mod 3; each owns one P pair).
//   TMEM per CTA: 4 S tiles (4 x 64 columns) + O (2 x 128 columns) = 512 columns.
//   Shared memory per CTA (226 KB): Q 40 KB, 2 K stages x 40 KB, a ring of 3 V units x 16 KB (one unit = one 256-dim
//   half of the output for the 128 tokens of a tile), 3 P pairs (hi, lo) x 2 x 8 KB (also the epilogue's staging), k-scale
//   ring 3 KB, barriers / row state / log2 and reciprocal scales 7 KB.
#include <string.h>

#include <type_traits>

#include "mla_common.cuh"

namespace b200 {
namespace pairk {

constexpr int kKStages = 2;
constexpr int kVUnits = 3;                  // V ring: unit u = 2 * tile + d-half, slot u % kVUnits
constexpr int kGroups = 3;                  // softmax groups of four warps; tile g belongs to group g % kGroups
constexpr int kKsStages = 6, kKsAheadMax = 4;  // k-scale ring: 512 B per tile, loaded ahead of the softmax
constexpr int kQBytes = 64 * 640;          // 4 nope chunks [64 rows x 128 B] + rope [64 x 128 B]
constexpr int kKStageBytes = 64 * 640;     // this CTA's page of the tile, same layout
constexpr int kVUnitBytes = 128 * 128;       // [128 tokens][128 B]: this CTA's 128 of the 256 output dims of one d-half
constexpr int kPBytes = 64 * 128;          // [64 q rows][128 tokens] e4m3, K-major SW128
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kKStages * kKStageBytes;
constexpr int kOffP = kOffV + kVUnits * kVUnitBytes;
constexpr int kPPairs = kGroups;                     // P tiles come in (hi, lo) pairs, one pair per softmax group
constexpr int kOffKs = kOffP + kPPairs * 2 * kPBytes;  // k-scale ring: [kKsStages][128] f32; a group's P pair (16 KB)
                                                     // doubles as its epilogue staging tile
constexpr int kOffSmall = kOffKs + kKsStages * 512;
constexpr int kSmallBytes = 7168;
constexpr int kTotal = kOffSmall + kSmallBytes;
constexpr int kSmThreads = 128 * kGroups;            // all softmax groups
constexpr int kSBufs = 4;                            // S tiles in TMEM: Q.K^T may run kSBufs - 1 tiles ahead of P.V
constexpr uint32_t kColS = 0, kColO = kSBufs * 64, kTmemCols = 512;
static_assert(kColO + 256 <= kTmemCols, "TMEM budget");
static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kOffP % 1024 == 0 && kOffKs % 1024 == 0, "swizzle alignment");

struct Small {
  uint64_t k_full[kKStages], k_empty[kKStages];
  uint64_t v_land[kVUnits], v_empty[kVUnits];
  uint64_t ks_full[kKsStages], ks_empty[kKsStages];
  uint64_t q_full, q_empty;
  uint64_t s_full[kSBufs], s_free[kSBufs];     // s_free(g): P.V of tile g retired
  uint64_t s_done[kSBufs];                     // both CTAs' softmax groups are through with S tile g % kSBufs: Q.K^T may overwrite it
  uint64_t v_aux[2];                           // landing barrier of a request's last units while their rows past the end are zeroed
  uint64_t v_peer[kVUnits];                    // leader: the peer CTA's V unit has landed (forwarded by its relay warp)
  uint64_t p_full[kPPairs], p_free[kPPairs];   // per P pair (tile parity)
  uint64_t pv_done, o_ready, o_free;
  uint64_t p_relay[kPPairs], o_relay;  // non-leader CTA: local hand-off points forwarded to the leader by its warp 1
  uint64_t fin[2];               // tile g (buffer g & 1) is final: published, reference moves done
  uint32_t tmem_base;
  uint32_t version;              // bumped by every reference move
  uint32_t k_issued;             // tiles whose K page this CTA's producer has issued (orders the V re-read behind it)
  uint32_t trig[kGroups][3];
  int merge_ticket;              // split request: how many of its parts had reached their epilogue before this one
  float st_ref[64];              // per-row reference exponent R (log2): O and L are in units of 2^R / 448
  float lred[2][2 * kGroups][64];  // row-sum exchange, double-buffered by request parity
  alignas(16) float lkrk[kGroups][256];  // per softmax group: log2(k_scale) [128], 1 / k_scale [128] of the tile it works on
};
static_assert(sizeof(Small) <= kSmallBytes, "Small too large");
static_assert(kTotal + 1024 <= 232448, "over the 227 KB shared-memory limit");

// Pipeline trace of part 0 (diagnostic): [rank][role: 0 producer, 1 mma, 2 softmax warp 4, 3 epilogue (indexed by
// request)][tile < 64][event < 4]
constexpr int kTraceTiles = 64;
constexpr int kTraceRoles = 5;
__device__ unsigned long long g_trace[2 * kTraceRoles * kTraceTiles * 4];
__device__ unsigned long long g_part_cycles[2 * 128 * 4];  // [part < 128][rank]{globaltimer ns at entry, at exit, clock64 at entry, at exit}
// "Chaos" mode of the diagnostic instantiation (B200_MLA_TRACE bit 2): every trace site -- 20 of them over the producer,
// the MMA issuers, the softmax passes, the hand-offs and the epilogue -- delays its warp by a pseudo-random 0-4 us, so
// warps, groups and the two CTAs of a pair drift against each other by whole tiles.  The protocol must not care: the
// results have to stay bit-identical (tests/test_gpu_parity.py::test_mla_pair_is_timing_independent).
__constant__ int g_chaos;  // (constant bank: a trace site must not cost an L2 round trip)
// TR = false (the production instantiation) compiles every trace site out
template <bool TR>
__device__ __forceinline__ void trace(bool on, uint32_t rank, int role, uint32_t g, int ev) {
  if constexpr (TR) {
    // one bit per site: 4 * role + event for the pipeline's sites, 20-25 for the request-boundary detail sites (role 3,
    // index >= 8), so that a perturbation that changes results can be pinned to the hand-off it disturbs
    // (B200_MLA_CHAOS_SITES)
    const int site_bit = (role == 3 && g >= 8u) ? (g >= 16u ? 24 + ev : 20 + ev) : 4 * role + ev;
    if ((g_chaos >> site_bit) & 1) {  // (the same delay for the whole warp: the trace sites are warp-converged code in front of aligned barriers)
      const uint32_t c = __shfl_sync(0xffffffffu, (uint32_t)clock64(), 0);
      uint32_t h = c * 2654435761u ^ ((threadIdx.x >> 5) * 40503u + blockIdx.z * 9176u + role * 131u + ev * 31u + g);
      h ^= h >> 13, h *= 0x5bd1e995u, h ^= h >> 15;
      if ((h & 3u) == 0) __nanosleep((h >> 4) & 0xfffu);
    }
    if (on && g < (uint32_t)kTraceTiles) g_trace[((rank * kTraceRoles + role) * kTraceTiles + g) * 4 + ev] = clock64();
  }
}

// Token rows past the end of a request meet P = 0 in P.V: make V there finite (0 * NaN).  Reference: fill_oob_V,
// splitkv_mla.cu:698-721.  Called by the V producer warp of each CTA, the one warp that sees every unit of its ring in
// order: a unit with such rows lands on an auxiliary barrier, the producer zeroes the rows and only then completes the
// unit's phase of v_land, so the consumers (P.V issuer, relay) see a finished unit and know nothing of this.  (The
// softmax groups used to do it in front of their publish.  They run up to three tiles ahead of P.V, though: with unit
// u - 3 of the same slot still in flight, their parity wait for unit u returned on the slot's older phase and the zeros
// went into the rows of unit u - 3 under its TMA -- found by the timing-perturbation test with delays in the V
// producer.)  At most 31 rows, one per lane: rows [valid_in_tile, nk * 32) of [128 tokens][128 B].
__device__ __forceinline__ void zero_v_rows_past_end(uint8_t* unit, int valid_in_tile, int nk, int lane) {
  const int r = valid_in_tile + lane;
  if (r < nk * 32) {
    uint4* vr = reinterpret_cast<uint4*>(unit + r * 128);
#pragma unroll
    for (int k = 0; k < 8; ++k) vr[k] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async_smem();
}

// O *= fac for this thread's TMEM lane (256 columns): a row reference moved
__device__ __forceinline__ void rescale_o(uint32_t lane_base, float fac) {
#pragma unroll 1
  for (int c8 = 0; c8 < 8; ++c8) {
    float o[32];
    const uint32_t a = lane_base + c8 * 32;
    tmem_ld<32>(a, o);
    tmem_wait_ld();
#pragma unroll
    for (int k = 0; k < 32; ++k) o[k] *= fac;
    tmem_st<32>(a, o);
  }
  tmem_wait_st();
}

// Register-resident window of one block-table row for the producer warp (warp-uniform use): lane l holds entries
// base + l and base + 32 + l.  With ~230 KB of shared memory per CTA there is no L1 left, so every __ldg of the block
// table is an L2 round trip (300-600 cycles under load); three of them per tile in the producer's dependency chain were
// the pipeline's critical path.  Entries are fetched 16+ tiles ahead instead and read with a shuffle.
struct BtWindow {
  const int32_t* row;
  int base, limit, cur, nxt;
  __device__ __forceinline__ void init(const int32_t* r, int first, int lim, int lane) {
    row = r, base = first, limit = lim;
    cur = (base + lane < limit) ? __ldg(row + base + lane) : 0;
    nxt = (base + 32 + lane < limit) ? __ldg(row + base + 32 + lane) : 0;
  }
  // window [0, 64) of a row whose entries were fetched while the previous request ran
  __device__ __forceinline__ void adopt(const int32_t* r, int lim, int c, int n) {
    row = r, base = 0, limit = lim, cur = c, nxt = n;
  }
  // call with the lowest block index still needed; keeps [lowest, lowest + 32) inside the window
  __device__ __forceinline__ void advance_to(int lowest, int lane) {
    while (lowest >= base + 32) {
      cur = nxt;
      base += 32;
      nxt = (base + 32 + lane < limit) ? __ldg(row + base + 32 + lane) : 0;
    }
  }
  __device__ __forceinline__ int get(int blk) const {
    const int i = blk - base;
    const int a = __shfl_sync(0xffffffffu, cur, i & 31), c = __shfl_sync(0xffffffffu, nxt, i & 31);
    return i < 32 ? a : c;
  }
};


template <bool TR, bool DEFER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128 + kSmThreads, 1)
    mla_decode_pair_kernel(const __grid_constant__ MlaParams p) {
  extern __shared__ uint8_t smem_raw[];
  // round up to 1024 B by pointer arithmetic (an integer round trip would demote every access to generic LD / ST)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* s_q = smem + kOffQ;
  uint8_t* s_k = smem + kOffK;
  uint8_t* s_v = smem + kOffV;
  uint8_t* s_p = smem + kOffP;
  float* s_ks = reinterpret_cast<float*>(smem + kOffKs);
  Small& sm = *reinterpret_cast<Small*>(smem + kOffSmall);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t rank = cluster_ctarank();
  const int part = blockIdx.z;
  const int pair_row0 = blockIdx.y * 128;
  const int q_row0 = pair_row0 + 64 * (int)rank;
  const int n_valid_q = max(0, min(64, p.M - q_row0));
  const bool tr = TR && part == 0 && blockIdx.y == 0 && lane == 0;
  if (TR && tid == 0 && part < 128 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_part_cycles[(part * 2 + rank) * 4] = t;
    g_part_cycles[(part * 2 + rank) * 4 + 2] = clock64();
  }

  // Launched as a programmatic dependent of the kernel in front of it (the prologue that writes q and the new K rows):
  // everything up to griddepcontrol.wait touches no global data and overlaps that kernel's tail.
  if (tid == 0) {
    for (int i = 0; i < kKStages; ++i) {
      mbar_init(&sm.k_full[i], 2);  // one expect_tx arrival per CTA of the pair (leader's copy is the live one)
      mbar_init(&sm.k_empty[i], 1);
    }
    for (int i = 0; i < kVUnits; ++i) {
      mbar_init(&sm.v_land[i], 1);
      mbar_init(&sm.v_empty[i], 1);
    }
    for (int i = 0; i < kKsStages; ++i) {
      mbar_init(&sm.ks_full[i], 1);
      mbar_init(&sm.ks_empty[i], 4);  // the four warps of the softmax group that serves the tile
    }
    mbar_init(&sm.q_full, 2);
    mbar_init(&sm.q_empty, 1);
    for (int i = 0; i < kSBufs; ++i) {
      mbar_init(&sm.s_full[i], 1);
      mbar_init(&sm.s_free[i], 1);
      mbar_init(&sm.s_done[i], 5);   // the four warps of the leader's softmax group + the peer's relay
    }
    for (int i = 0; i < kVUnits; ++i) mbar_init(&sm.v_peer[i], 1);
    for (int i = 0; i < 2; ++i) mbar_init(&sm.v_aux[i], 1);
    for (int i = 0; i < kPPairs; ++i) {
      mbar_init(&sm.p_full[i], 5);   // the four warps of the leader's softmax group + the peer's relay
      mbar_init(&sm.p_relay[i], 4);
      mbar_init(&sm.p_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&sm.fin[i], 4);
    mbar_init(&sm.pv_done, 1);
    mbar_init(&sm.o_ready, 1);
    mbar_init(&sm.o_free, 4 * kGroups + 1);   // the leader's softmax warps (all groups write a request out) + the peer's relay
    mbar_init(&sm.o_relay, 4 * kGroups);
    for (int i = 0; i < 3 * kGroups; ++i) (&sm.trig[0][0])[i] = 0;
    sm.version = 0;
    sm.k_issued = 0;
    fence_mbar_init();
    tma_prefetch_desc(&p.tm_q_nope);
    tma_prefetch_desc(&p.tm_q_rope);
    tma_prefetch_desc(&p.tm_k_nope);
    tma_prefetch_desc(&p.tm_k_rope);
  }
  __syncwarp();
  if (warp == 1) tmem_alloc_pair(&sm.tmem_base, kTmemCols);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised before anything is credited to them
  tc_fence_after();
  const uint32_t tmem = sm.tmem_base;

  pdl_wait_prior_grid();
  const int32_t* meta = p.meta + part * B200_MLA_METADATA_INTS;
  const int begin_idx = __ldg(meta + 0), begin_seqlen = __ldg(meta + 1), end_idx = __ldg(meta + 2),
            end_seqlen = __ldg(meta + 3), begin_n_split = __ldg(meta + 4);
  // a kernel behind this one that is launched with programmatic stream serialization may be scheduled as SMs drain
  // (its own griddepcontrol.wait orders its reads)
  pdl_launch_dependents();
  if (begin_idx >= p.batch) {  // splitkv_mla.cu:1226 (both CTAs of the pair take the same exit)
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_pair(tmem, kTmemCols);
    return;
  }

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    const uint32_t q_full_l = mapa_u32(smem_u32(&sm.q_full), 0);
    uint32_t k_full_l[kKStages];
#pragma unroll
    for (int i = 0; i < kKStages; ++i) k_full_l[i] = mapa_u32(smem_u32(&sm.k_full[i]), 0);

    uint32_t g = 0, seg_i = 0;
    const int pf = min(p.pf_tiles, 14);  // the block-table window holds 32 entries past the current tile
    const int ks_ahead = min(p.ks_ahead, kKsAheadMax);
    // Everything a request's first loads depend on is fetched while the previous request runs (its lengths, its split
    // index, the first 64 entries of its block-table row): with no L1, the four dependent L2 round trips otherwise sit
    // between the last K page of one request and the first of the next.
    int cur_sl = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);
    int pre_cur = 0, pre_nxt = 0;
    bool have_pre = false;
    int ks_done = 0;  // tiles of this request whose k scales were already issued from the previous request's tail
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment_from(p, b, cur_sl, cur_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                           begin_n_split, pair_row0, 128);
      int nxt_sl = 0, nxt_ns = 0;
      if (b < end_idx) {
        nxt_sl = __ldg(p.seqlens + b + 1);
        nxt_ns = __ldg(p.num_splits + b + 1);
      }
      if (sg.n_tiles == 0) {
        cur_sl = nxt_sl, cur_ns = nxt_ns, have_pre = false, ks_done = 0;
        continue;
      }
      const int32_t* bt = p.block_table + (int64_t)b * p.block_table_stride;
      BtWindow bw;
      if (have_pre) bw.adopt(bt, sg.end_block, pre_cur, pre_nxt);
      else bw.init(bt, sg.start_block, sg.end_block, lane);
      // the request after this one (same part, so it starts at block 0): filled in below, once this request's first
      // loads are out
      const int32_t* bt_next = nullptr;
      int nxt_blocks = 0, nxt_tiles = 0, nxt_req = 0, nxt_req2 = 0;

      // this CTA's page of tile t -> K stage (credits the leader's barrier)
      auto load_k = [&](int t, uint32_t gt) {
        const uint32_t st = gt % kKStages;
        mbar_wait(&sm.k_empty[st], ((gt / kKStages) & 1) ^ 1);
        const int blk = sg.start_block + 2 * t + (int)rank;
        const bool exists = blk < sg.end_block;
        const int tok0 = exists ? bw.get(blk) * kPage : 0;
        // L2 prefetch `pf` tiles ahead (past the end of this request: the first tiles of the next one)
        int tok_pf = -1;
        {
          const int tp = t + pf;
          if (tp < sg.n_tiles) {
            const int bp = sg.start_block + 2 * tp + (int)rank;
            if (bp < sg.end_block) tok_pf = bw.get(bp) * kPage;
          } else if (bt_next != nullptr) {
            const int bp = 2 * (tp - sg.n_tiles) + (int)rank;
            if (bp < nxt_blocks && bp < 32) tok_pf = __shfl_sync(0xffffffffu, nxt_req, bp) * kPage;
          }
        }
        uint8_t* stage = s_k + st * kKStageBytes;
        trace<TR>(tr, rank, 0, gt, 0);
        if (elect_one()) {
          mbar_arrive_expect_tx_cluster(k_full_l[st], exists ? (uint32_t)kKStageBytes : 0u);
          if (exists) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_load_2d_pair(stage + c * 8192, &p.tm_k_nope, k_full_l[st], c * 128, tok0);
            tma_load_2d_pair(stage + 32768, &p.tm_k_rope, k_full_l[st], 0, tok0);
          }
          if (tok_pf >= 0) {  // warm L2 for a page `pf` tiles ahead
            if (p.pf_mode == 0) {
#pragma unroll
              for (int c = 0; c < 4; ++c) tma_prefetch_2d(&p.tm_k_nope, c * 128, tok_pf);
              tma_prefetch_2d(&p.tm_k_rope, 0, tok_pf);
            } else if (p.pf_mode == 2) {  // the page's 32 KB of latent and 8 KB of rope are contiguous: two bulk prefetches
              bulk_prefetch_l2(p.k_lora_base + (int64_t)tok_pf * 512, 64 * 512);
              bulk_prefetch_l2(p.k_rope_base + (int64_t)tok_pf * 128, 64 * 128);
            }
          }
        }
        __syncwarp();
        if (lane == 0) st_shared_volatile_u32(&sm.k_issued, gt + 1);
      };
      // k scales of both pages of tile t -> scale ring.  Their HBM latency (a separate array: no L2 warm-up by the K
      // stream) is paid kKsAhead tiles before the softmax needs them.
      auto load_ks = [&](int t, uint32_t gt) {
        const uint32_t st = gt % kKsStages;
        mbar_wait(&sm.ks_empty[st], ((gt / kKsStages) & 1) ^ 1);
        const int blk0 = sg.start_block + 2 * t;
        const int npages = min(2, sg.end_block - blk0);
        const int tok0 = bw.get(blk0) * kPage;
        const int tok1 = npages > 1 ? bw.get(blk0 + 1) * kPage : 0;
        float* ks_dst = s_ks + st * 128;
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.ks_full[st], (uint32_t)npages * 256);
          bulk_load_1d(ks_dst, p.k_scale + tok0, 256, &sm.ks_full[st]);
          if (npages > 1) bulk_load_1d(ks_dst + 64, p.k_scale + tok1, 256, &sm.ks_full[st]);
        }
        __syncwarp();
      };
      // the same for tile tn of the NEXT request: the scale ring runs across the request boundary, so the first tiles of a
      // request do not wait for an HBM round trip that starts at the boundary
      auto load_ks_next = [&](int tn, uint32_t gt) {
        const uint32_t st = gt % kKsStages;
        mbar_wait(&sm.ks_empty[st], ((gt / kKsStages) & 1) ^ 1);
        const int blk0 = 2 * tn;
        const int npages = min(2, nxt_blocks - blk0);
        const int e0 = __shfl_sync(0xffffffffu, nxt_req, blk0 & 31), e1 = __shfl_sync(0xffffffffu, nxt_req, (blk0 + 1) & 31);
        float* ks_dst = s_ks + st * 128;
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.ks_full[st], (uint32_t)npages * 256);
          bulk_load_1d(ks_dst, p.k_scale + e0 * kPage, 256, &sm.ks_full[st]);
          if (npages > 1) bulk_load_1d(ks_dst + 64, p.k_scale + e1 * kPage, 256, &sm.ks_full[st]);
        }
        __syncwarp();
      };

      // The first K page (+ its scales) and Q are what the first Q.K^T waits for.  K goes first: its stage frees one
      // tile before the previous request's last Q.K^T retires (q_empty), and at a request boundary the K page comes from
      // HBM while Q was prefetched into L2.
      load_k(0, g);
      if (ks_done == 0) {
        load_ks(0, g);
        ks_done = 1;
      }
      // Q rows of this CTA (rows past the request are masked by the consumers)
      mbar_wait(&sm.q_empty, (seg_i & 1) ^ 1);
      ++seg_i;
      const int qrow = b * p.M + q_row0;
      if (elect_one()) {
        mbar_arrive_expect_tx_cluster(q_full_l, kQBytes);
#pragma unroll
        for (int c = 0; c < 4; ++c) tma_load_2d_pair(s_q + c * 8192, &p.tm_q_nope, q_full_l, c * 128, qrow);
        tma_load_2d_pair(s_q + 32768, &p.tm_q_rope, q_full_l, 0, qrow);
      }
      __syncwarp();
      if (b < end_idx) {
        const Segment sgn = make_segment_from(p, b + 1, nxt_sl, nxt_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                              begin_n_split, pair_row0, 128);
        nxt_blocks = sgn.end_block;
        nxt_tiles = sgn.n_tiles;
        bt_next = p.block_table + (int64_t)(b + 1) * p.block_table_stride;
        nxt_req = lane < nxt_blocks ? __ldg(bt_next + lane) : 0;  // its first 16 tiles
        nxt_req2 = 32 + lane < nxt_blocks ? __ldg(bt_next + 32 + lane) : 0;
      }
      int ks_next = 0;  // k-scale tiles of the next request issued so far
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        bw.advance_to(sg.start_block + 2 * t, lane);
        if (t + 1 < sg.n_tiles) load_k(t + 1, g + 1);
        if (t == 0 && seg_i == 1 && p.pf_mode != 1) {  // first request of this part: tiles the steady-state prefetch (t + pf)
                                                      // will not reach, behind the demand loads of tiles 0 and 1
          for (int tp = kKStages; tp < pf && tp < sg.n_tiles; ++tp) {
            const int bp = sg.start_block + 2 * tp + (int)rank;
            if (bp < sg.end_block) {
              const int tk = bw.get(bp) * kPage;
              if (elect_one()) {
                if (p.pf_mode == 0) {
#pragma unroll
                  for (int c = 0; c < 4; ++c) tma_prefetch_2d(&p.tm_k_nope, c * 128, tk);
                  tma_prefetch_2d(&p.tm_k_rope, 0, tk);
                } else {
                  bulk_prefetch_l2(p.k_lora_base + (int64_t)tk * 512, 64 * 512);
                  bulk_prefetch_l2(p.k_rope_base + (int64_t)tk * 128, 64 * 128);
                }
              }
              __syncwarp();
            }
          }
        }
        if (t == 0)
          for (int i = ks_done; i < ks_ahead && i < sg.n_tiles; ++i) load_ks(i, g + i);
        // next request's Q rows of this CTA -> L2, a few tiles before they are loaded: any earlier and the KV stream
        // (about 4 MB per tile time chip-wide) has pushed them out of L2 again by the time the request starts
        if (b < end_idx && t == (sg.n_tiles > p.q_pf_tiles ? sg.n_tiles - p.q_pf_tiles : 0)) {
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 4; ++c) tma_prefetch_2d(&p.tm_q_nope, c * 128, qrow + p.M);
            tma_prefetch_2d(&p.tm_q_rope, 0, qrow + p.M);
          }
          __syncwarp();
        }
        if (t + ks_ahead < sg.n_tiles) load_ks(t + ks_ahead, g + ks_ahead);
        // scales of the next request's first tiles, kKsAhead tiles ahead as well (in tile order: behind this request's)
        while (ks_next < nxt_tiles && ks_next < 16 && ks_next <= t + ks_ahead - sg.n_tiles) {
          load_ks_next(ks_next, g + (uint32_t)(sg.n_tiles - t + ks_next));
          ++ks_next;
        }
      }
      cur_sl = nxt_sl, cur_ns = nxt_ns;
      pre_cur = nxt_req, pre_nxt = nxt_req2, have_pre = bt_next != nullptr;
      ks_done = ks_next;
    }
  } else if (TR && (p.trace & 64) && warp == 3 && rank == 1 && part == 0 && blockIdx.y == 0) {
    // diagnostic (B200_MLA_TRACE bit 6, tools/trace_mla.py; needs at least 40 tiles in part 0 -- a passive observer of
    // barrier parities must not fall behind): when the second V unit of each tile lands in the non-leader CTA
    for (uint32_t g = 0; g < 40; ++g) {
      mbar_wait(&sm.v_land[(2 * g + 1) % kVUnits], ((2 * g + 1) / kVUnits) & 1);
      trace<TR>(lane == 0, rank, 0, g, 3);
    }
  } else if (warp == 2) {
    // ============================== V producer (both CTAs) ==============================
    // Its own warp: a V unit frees when the P.V products that read it retire, which must not hold back the K stream
    // (Q.K^T runs ahead of P.V).  A unit is one 256-dim half of the output (d-half j) for the 128 tokens of a tile:
    // unit u = 2 * tile + j lives in slot u % 3, so the ring holds a tile and a half and the second half of tile g + 1
    // is refilled while P.V of tile g still reads the other half.
    const uint64_t pol_last_use = policy_evict_first();
    uint32_t g = 0, n_tails = 0;  // n_tails: units pairs that went through v_aux (its phase count)
    int cur_sl = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);  // fetched one request ahead
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment_from(p, b, cur_sl, cur_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                           begin_n_split, pair_row0, 128);
      if (b < end_idx) {
        cur_sl = __ldg(p.seqlens + b + 1);
        cur_ns = __ldg(p.num_splits + b + 1);
      }
      if (sg.n_tiles == 0) continue;
      const int32_t* bt = p.block_table + (int64_t)b * p.block_table_stride;
      BtWindow bw;
      bw.init(bt, sg.start_block, sg.end_block, lane);
      // both pages of tile t, this CTA's 128-dim slice of d-half j -> V unit (local barrier); these bytes were just
      // streamed through L2 as K by this CTA or its peer
      auto load_v = [&](int t, uint32_t gt) {
        trace<TR>(tr, rank, 0, gt, 1);
        B200_BEACON(1, gt);
        const int blk0 = sg.start_block + 2 * t;
        const int npages = min(2, sg.end_block - blk0);
        const int tok0 = bw.get(blk0) * kPage;
        const int tok1 = npages > 1 ? bw.get(blk0 + 1) * kPage : 0;
        // rows past the request's end inside the 32-token blocks P.V reads (only a request's last tile has any)
        const int valid_in_tile = sg.valid_end - (sg.start_block + 2 * t) * kPage;
        const int nk = (min(kTileTokens, valid_in_tile) + 31) / 32;
        const bool zero_tail = valid_in_tile < nk * 32 && !(TR && (p.trace & 8));
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const uint32_t u = 2 * gt + (uint32_t)j, st = u % kVUnits;
          mbar_wait(&sm.v_empty[st], ((u / kVUnits) & 1) ^ 1);
          if (j == 0) {
            // behind this tile's K issue: the same bytes then come out of L2 (or merge with the K fetch in flight)
            // instead of competing with the K and Q loads for HBM at a request's start
            for (uint32_t spins = 0; !(TR && (p.trace & 32)) && ld_shared_volatile_u32(&sm.k_issued) <= gt; ++spins) {
              if (spins > B200_WAIT_SPIN_LIMIT) __trap();
              __nanosleep(64);
            }
            trace<TR>(tr, rank, 0, gt, 2);
          }
          uint8_t* unit = s_v + st * kVUnitBytes;
          if (TR && (p.trace & 8)) {  // diagnostic (wrong results): no V traffic at all
            if (elect_one()) mbar_arrive(&sm.v_land[st]);
            __syncwarp();
            continue;
          }
          if (elect_one()) {
            uint64_t* land = zero_tail ? &sm.v_aux[j] : &sm.v_land[st];
            mbar_arrive_expect_tx(land, (uint32_t)npages * 8192);
            tma_load_2d_hint(unit, &p.tm_k_nope, land, 256 * j + 128 * (int)rank, tok0, pol_last_use);
            if (npages > 1) tma_load_2d_hint(unit + 8192, &p.tm_k_nope, land, 256 * j + 128 * (int)rank, tok1, pol_last_use);
          }
          __syncwarp();
        }
        if (zero_tail) {  // (the next unit's slot frees no earlier than P.V of this tile's predecessor: nothing waits on this)
#pragma unroll 1
          for (int j = 0; j < 2; ++j) {
            const uint32_t u = 2 * gt + (uint32_t)j, st = u % kVUnits;
            mbar_wait(&sm.v_aux[j], n_tails & 1);
            zero_v_rows_past_end(s_v + st * kVUnitBytes, valid_in_tile, nk, lane);
            __syncwarp();
            if (elect_one()) mbar_arrive(&sm.v_land[st]);  // the unit's phase of v_land: landed and finished
            __syncwarp();
          }
          ++n_tails;
        }
      };
      for (int t = 0; t < sg.n_tiles; ++t, ++g) {
        bw.advance_to(sg.start_block + 2 * t, lane);
        load_v(t, g);
      }
    }
  } else if ((warp == 1 || warp == 3) && rank == 0) {
    // ============================== UMMA issuers (leader CTA) ==============================
    // warp 1 issues every Q.K^T, warp 3 every P.V: two independent instruction streams into the one in-order tensor
    // pipe, so a Q.K^T waiting for its K page never holds back a P.V whose P is published (and vice versa; a single
    // issuer needs either a fixed interleave or polling, and both cost throughput).  Q.K^T of tile g waits until the
    // softmax groups of both CTAs are through with tile g - kSBufs (s_done: S tile g % kSBufs is dead as soon as its P
    // is published, long before that P.V retires).
    constexpr uint32_t idesc_qk8 = make_idesc(FMT_E4M3, FMT_E4M3, 0, 0, 128, 128);
    constexpr uint32_t idesc_qk16 = make_idesc(FMT_BF16, FMT_BF16, 0, 0, 128, 128);
    constexpr uint32_t idesc_pv = make_idesc(FMT_E4M3, FMT_E4M3, 0, 1, 128, 256);
    const uint32_t q_addr = smem_u32(s_q), k_addr = smem_u32(s_k), v_addr = smem_u32(s_v), p_addr = smem_u32(s_p);
    const uint64_t d_kmaj = make_smem_desc(0, 0, 1024, SWZ_128B);       // K-major, 128-B swizzle
    const uint64_t d_vmn = make_smem_desc(0, 16384, 1024, SWZ_128B);    // V: MN-major view of the token rows
    auto with_addr = [](uint64_t d, uint32_t a) { return d | (uint64_t)((a >> 4) & 0x3FFF); };
    uint32_t g = 0, seg = 0;
    int cur_sl = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);  // fetched one request ahead
    for (int b = begin_idx; b <= end_idx; ++b) {
      const Segment sg = make_segment_from(p, b, cur_sl, cur_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                           begin_n_split, pair_row0, 128);
      if (b < end_idx) {
        cur_sl = __ldg(p.seqlens + b + 1);
        cur_ns = __ldg(p.num_splits + b + 1);
      }
      if (sg.n_tiles == 0) continue;
      if (warp == 1) {
        mbar_wait_cluster(&sm.q_full, seg & 1);
        for (int t = 0; t < sg.n_tiles; ++t, ++g) {
          const uint32_t kst = g % kKStages, sb = g % kSBufs;
          if (g >= (uint32_t)kSBufs) mbar_wait_cluster(&sm.s_done[sb], ((g / kSBufs) & 1) ^ 1);  // S tile g - kSBufs consumed
          mbar_wait(&sm.ks_full[g % kKsStages], (g / kKsStages) & 1);  // s_full then vouches for the k scales too
          mbar_wait_cluster(&sm.k_full[kst], (g / kKStages) & 1);
          trace<TR>(tr, rank, 1, g, 0);
          B200_BEACON(2, g);
          tc_fence_after();
          const uint32_t stage = k_addr + kst * kKStageBytes;
          const uint32_t d = tmem + kColS + sb * 64;
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
              for (int s = 0; s < 4; ++s)
                umma_f8_pair(d, with_addr(d_kmaj, q_addr + c * 8192 + s * 32), with_addr(d_kmaj, stage + c * 8192 + s * 32),
                             idesc_qk8, (c | s) != 0);
#pragma unroll
            for (int s = 0; s < 4; ++s)
              umma_f16_pair(d, with_addr(d_kmaj, q_addr + 32768 + s * 32), with_addr(d_kmaj, stage + 32768 + s * 32),
                            idesc_qk16, 1);
            umma_commit_pair(&sm.s_full[sb], 3);
            umma_commit_pair(&sm.k_empty[kst], 3);
            if (t == sg.n_tiles - 1) umma_commit_pair(&sm.q_empty, 3);  // Q smem reusable once these retire
          }
          __syncwarp();
          trace<TR>(tr, rank, 1, g, 1);
        }
      } else {
        for (int t = 0; t < sg.n_tiles; ++t, ++g) {
          const uint32_t pp = g % kPPairs;
          mbar_wait_cluster(&sm.p_full[pp], (g / kPPairs) & 1);  // P written in both CTAs
          if (t == 0) trace<TR>(tr, rank, 3, 16 + seg, 0);
          if (t == 0 && seg > 0) mbar_wait_cluster(&sm.o_free, (seg - 1) & 1);
          if (t == 0) trace<TR>(tr, rank, 3, 16 + seg, 1);
          trace<TR>(tr, rank, 1, g, 2);
          B200_BEACON(3, g);
          tc_fence_after();
          const int tile_start = (sg.start_block + 2 * t) * kPage;
          const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 31) / 32;
#pragma unroll
          for (int j = 0; j < 2; ++j) {  // d-half by d-half: the first V unit is released while the second is read
            const uint32_t u = 2 * g + (uint32_t)j, vst = u % kVUnits;
            mbar_wait(&sm.v_land[vst], (u / kVUnits) & 1);          // V unit landed here
            mbar_wait_cluster(&sm.v_peer[vst], (u / kVUnits) & 1);  // ... and in the peer
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
              for (int hl = 0; hl < 2; ++hl)  // P = P_hi + P_lo
#pragma unroll
                for (int s = 0; s < 4; ++s)
                  if (s < nk)
                    umma_f8_pair(tmem + kColO + j * 128, with_addr(d_kmaj, p_addr + (2 * pp + hl) * kPBytes + s * 32),
                                 with_addr(d_vmn, v_addr + vst * kVUnitBytes + s * 4096), idesc_pv,
                                 !(t == 0 && s == 0 && hl == 0));
              umma_commit_pair(&sm.v_empty[vst], 3);
            }
            __syncwarp();
          }
          if (elect_one()) {
            umma_commit_pair(&sm.p_free[pp], 3);
            // pv_done in front of s_free: a reference move waits for s_free of tile g - 2 to know that pv_done is at most
            // one phase behind its own parity wait; with the two arrivals in this order, seeing the former implies the
            // latter
            umma_commit_pair(&sm.pv_done, 3);
            umma_commit_pair(&sm.s_free[g % kSBufs], 3);
            if (t == sg.n_tiles - 1) umma_commit_pair(&sm.o_ready, 3);  // one completion per request: the epilogue's gate
          }
          __syncwarp();
          trace<TR>(tr, rank, 1, g, 3);
        }
      }
      ++seg;
    }
  } else if (warp == 1) {
    {
      // ============================== relay (non-leader CTA) ==============================
      // This CTA's softmax warps hand P / O over on LOCAL barriers (ordered with their shared-memory stores); once a
      // phase completes here -- the stores are then performed in this SM's shared memory -- one remote arrive tells the
      // leader.  A direct remote arrive from the softmax warps would need a cluster-scope release (a cluster-wide
      // membar per tile, measured at a third of their time); with CTA-scope release the arrive overtakes the stores.
      uint32_t p_full_l[kPPairs];
#pragma unroll
      for (int i = 0; i < kPPairs; ++i) p_full_l[i] = mapa_u32(smem_u32(&sm.p_full[i]), 0);
      const uint32_t o_free_l = mapa_u32(smem_u32(&sm.o_free), 0);
      uint32_t s_done_l[kSBufs], v_peer_l[kVUnits];
#pragma unroll
      for (int i = 0; i < kSBufs; ++i) s_done_l[i] = mapa_u32(smem_u32(&sm.s_done[i]), 0);
#pragma unroll
      for (int i = 0; i < kVUnits; ++i) v_peer_l[i] = mapa_u32(smem_u32(&sm.v_peer[i]), 0);
      uint32_t g = 0, seg = 0;
      int cur_sl = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);  // fetched one request ahead
      for (int b = begin_idx; b <= end_idx; ++b) {
        const Segment sg = make_segment_from(p, b, cur_sl, cur_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                             begin_n_split, pair_row0, 128);
        if (b < end_idx) {
          cur_sl = __ldg(p.seqlens + b + 1);
          cur_ns = __ldg(p.num_splits + b + 1);
        }
        if (sg.n_tiles == 0) continue;
        for (int t = 0; t < sg.n_tiles; ++t, ++g) {
          // per tile: the first V unit (frees early), then this CTA's P (its S tile is dead with it), then the second
          // V unit (the last to land: its slot frees when P.V of the previous tile's first half retires)
          const uint32_t u0 = 2 * g, u1 = 2 * g + 1;
          mbar_wait(&sm.v_land[u0 % kVUnits], (u0 / kVUnits) & 1);
          if (elect_one()) mbar_arrive_cluster(v_peer_l[u0 % kVUnits]);
          __syncwarp();
          B200_BEACON(4, g);
          mbar_wait(&sm.p_relay[g % kPPairs], (g / kPPairs) & 1);
          if (elect_one()) {
            mbar_arrive_cluster(s_done_l[g % kSBufs]);
            mbar_arrive_cluster(p_full_l[g % kPPairs]);
          }
          __syncwarp();
          mbar_wait(&sm.v_land[u1 % kVUnits], (u1 / kVUnits) & 1);
          if (elect_one()) mbar_arrive_cluster(v_peer_l[u1 % kVUnits]);
          __syncwarp();
        }
        B200_BEACON(4, 1000 + seg);
        mbar_wait(&sm.o_relay, seg & 1);
        if (elect_one()) mbar_arrive_cluster(o_free_l);
        __syncwarp();
        ++seg;
      }
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue (both CTAs) ==============================
    // Three groups of four warps take the tiles round-robin: group i (warps 4+4i .. 7+4i) serves the tiles with
    // g % 3 == i and owns P pair i.  A group's per-tile chain (barrier waits, two passes over S, vote, publish) then
    // overlaps the other groups' arithmetic instead of idling the SM: three softmax warps per scheduler.  A tile's P
    // depends on its predecessor only through the row reference R, which moves rarely ("trigger"), so the groups only
    // meet at a per-tile hand-off (sm.fin) that orders the moves: tile g is published after tile g-1 is final, and a
    // group that computed under a reference which has since moved (sm.version) recomputes its second pass.
    const int st_id = tid - 128;              // 0 .. 383 over the three groups
    const int grp = (warp - 4) >> 2;          // which tiles (g % 3) and which P pair this warp serves
    const int gtid = st_id & 127;             // thread within the group
    const uint32_t gbar = 2 + grp;            // the group's named barrier (128 threads)
    const int quad = warp & 3;                // TMEM lane quadrant of this warp
    const int tl = quad * 32 + lane;          // TMEM lane
    const int row = tl & 63;                  // q row within this CTA
    const int page = tl >> 6;                 // lanes 64-127 hold the second half of the N columns = the peer's page
    const int slot = page * kGroups + grp;    // this thread's slot in the row-sum exchange
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const bool row_valid = row < n_valid_q;
    const int grow = q_row0 + row;            // row within the request's M rows
    uint32_t g0 = 0, o_seg = 0;               // g0: global index of the request's first tile
    uint32_t n_mine = 0;                      // tiles this group has served (trigger-flag ring)
    // per-request scalars, fetched one request ahead (three dependent L2 round trips otherwise sit between a request's
    // epilogue and the next request's first tile)
    int cur_seqlen = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);
    const int ns_end = __ldg(p.num_splits + end_idx + 1);  // end of the last request's split range (merge decision)
    float cur_qs = row_valid ? __ldg(p.q_scale + (int64_t)begin_idx * p.M + grow) : 0.f;
    if (grp == 0 && page == 0) sm.st_ref[row] = kMInit;
    named_bar_sync(1, kSmThreads);
    uint32_t rq = 0;  // requests of this part seen so far (parity selects the row-sum exchange buffer)

    for (int b = begin_idx; b <= end_idx; ++b, ++rq) {
      const Segment sg = make_segment_from(p, b, cur_seqlen, cur_ns, begin_idx, begin_seqlen, end_idx, end_seqlen,
                                           begin_n_split, pair_row0, 128);
      // ---- per-request state ----
      const float f = cur_qs * p.scale_log2;
      if (b < end_idx) {  // next request's scalars: in flight while this request runs
        cur_seqlen = __ldg(p.seqlens + b + 1);
        cur_ns = __ldg(p.num_splits + b + 1);
        cur_qs = row_valid ? __ldg(p.q_scale + (int64_t)(b + 1) * p.M + grow) : 0.f;
      }
      int lim = 0;
      if (row_valid) lim = p.is_causal ? min(sg.seqlen - mask_len(p.s_q, p.h_q, p.M, grow), sg.end_block * kPage) : sg.valid_end;
      // The row state of this request was reset before the first request (below) or by the group that owns this
      // request's first tile, right after the previous request's last hand-off.  A thread's view starts invalid, so
      // its first tile loads the state (the other group may already have moved the reference).
      float rloc = kMInit, lsum = 0.f;  // this thread's view of the row reference R; lsum is in units of 2^rloc / 448
      uint32_t ver = 0xffffffffu;
      // bring this thread's view (rloc, lsum) up to the row's current reference
      auto sync_view = [&]() {
        // (lane 0's reading for the whole warp: what the warp does next -- recompute or not, the warp-wide votes -- must
        // not depend on one load instruction handing every lane the same value, e.g. under an instrumenting tool)
        const uint32_t v_now = __shfl_sync(0xffffffffu, ld_shared_volatile_u32(&sm.version), 0);
        if (v_now != ver) {
          const float r_new = ld_shared_volatile_f32(&sm.st_ref[row]);
          lsum *= exp2f(rloc - r_new);
          rloc = r_new;
          ver = v_now;
        }
      };

      // Pass A over this lane's 64 S columns: y_j = s_j f k_j + log2 k_j = log2 of the un-normalised P'_j, written back
      // over S in TMEM (masked columns as -inf), and ymax = max_j y_j (splitkv_mla.cu:400-432 computes the same maximum
      // as max_j exp2(..) k_j).  16 columns at a time in a compact loop with ONE copy of the body: the kernel's hot code
      // has to stay inside the instruction cache (a double-buffered body, twice the code, measured 7 % slower; the
      // TMEM load latency is covered by the other groups' warps).  The unmasked variant has no branches, so the 16
      // columns of a chunk are scheduled as independent chains.
      auto pass_max = [&](auto mask_tag, uint32_t s_addr, const float* ks_tile, const float* lk_tile, int tile_start) {
        constexpr bool need_mask = decltype(mask_tag)::value;
        const float2 f2 = make_float2(f, f);
        const float4* ksrc = reinterpret_cast<const float4*>(ks_tile + page * 64);
        const float4* lsrc = reinterpret_cast<const float4*>(lk_tile + page * 64);
        const int nvalid = lim - (tile_start + page * 64);  // columns of this lane below the mask border
        float m0 = -INFINITY, m1 = -INFINITY;
        auto chunk = [&](float* x, int c16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 k4 = ksrc[c16 * 4 + i], l4 = lsrc[c16 * 4 + i];
            if constexpr (!need_mask) {
              const float2 ya = fma2(make_float2(x[4 * i], x[4 * i + 1]), mul2(f2, make_float2(k4.x, k4.y)), make_float2(l4.x, l4.y));
              const float2 yb = fma2(make_float2(x[4 * i + 2], x[4 * i + 3]), mul2(f2, make_float2(k4.z, k4.w)), make_float2(l4.z, l4.w));
              m0 = fmaxf(m0, fmaxf(ya.x, ya.y));
              m1 = fmaxf(m1, fmaxf(yb.x, yb.y));
              x[4 * i] = ya.x, x[4 * i + 1] = ya.y, x[4 * i + 2] = yb.x, x[4 * i + 3] = yb.y;
            } else {
              const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float y = fmaf(x[4 * i + j], f * kk[j], ll[j]);
                // a select: S of a poisoned (NaN) padding row never propagates
                x[4 * i + j] = (c16 * 16 + 4 * i + j < nvalid) ? y : -INFINITY;
                m0 = fmaxf(m0, x[4 * i + j]);
              }
            }
          }
          tmem_st16(s_addr + c16 * 16, reinterpret_cast<const uint32_t*>(x));
        };
        // one copy of the 16-column body (half the code of the double-buffered form): the TMEM load latency of a chunk
        // is covered by the other softmax groups' warps on the same scheduler
        float xa[16];
#pragma unroll 1
        for (int c16 = 0; c16 < 4; ++c16) {
          tmem_ld16(s_addr + c16 * 16, reinterpret_cast<uint32_t*>(xa));
          tmem_wait_ld();
          chunk(xa, c16);
          tmem_wait_st();  // xa is the source of the store just issued
        }
        return fmaxf(m0, m1);
      };
      // Pass B: the page's codes code_j = e4m3(448 * 2^(y_j - ymax)) (:433-440), the page sum sum_j 448 * 2^(y_j - ymax)
      // / k_j (:426,434), and code_j * w as an e4m3 pair hi + lo straight into the two P tiles (K-major, 128-byte
      // swizzle; nobody reads them before this group arrives on p_full).  y comes back from TMEM; a masked column is
      // -inf there, i.e. code 0 and no contribution to the sum (its reciprocal scale was forced finite), so one variant
      // serves masked and unmasked tiles.
      auto pass_codes = [&](uint32_t s_addr, const float* rk_tile, float ymax, uint32_t w_h2, uint8_t* prow, float& acc) {
        const float c = (ymax > -INFINITY ? ymax : 0.f) - kLog2Fp8Max;
        const float2 nc2 = make_float2(-c, -c), one2 = make_float2(1.f, 1.f);
        float2 acc2 = make_float2(0.f, 0.f);
        const float4* rsrc = reinterpret_cast<const float4*>(rk_tile + page * 64);
        auto chunk = [&](const float* x, int c16) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 r4 = rsrc[c16 * 4 + i];
            const float2 aa = fma2(make_float2(x[4 * i], x[4 * i + 1]), one2, nc2);
            const float2 ab = fma2(make_float2(x[4 * i + 2], x[4 * i + 3]), one2, nc2);
            const float2 va = make_float2(ex2_approx(aa.x), ex2_approx(aa.y));
            const float2 vb = make_float2(ex2_approx(ab.x), ex2_approx(ab.y));
            acc2 = fma2(va, make_float2(r4.x, r4.y), acc2);
            acc2 = fma2(vb, make_float2(r4.z, r4.w), acc2);
            uint32_t h0, l0, h1, l1;
            split_hi_lo(cvt_e4m3x2(va.x, va.y), w_h2, h0, l0);
            split_hi_lo(cvt_e4m3x2(vb.x, vb.y), w_h2, h1, l1);
            hi[i] = h0 | (h1 << 16);
            lo[i] = l0 | (l1 << 16);
          }
          const int off = ((page * 4 + c16) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(prow + kPBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        };
        float xa[16];
#pragma unroll 1
        for (int c16 = 0; c16 < 4; ++c16) {
          tmem_ld16(s_addr + c16 * 16, reinterpret_cast<uint32_t*>(xa));
          tmem_wait_ld();
          chunk(xa, c16);
        }
        acc = acc2.x + acc2.y;
      };

      const int t_first = (grp + kGroups - (int)(g0 % kGroups)) % kGroups;  // this group's first tile of the request
      for (int t = t_first; t < sg.n_tiles; t += kGroups, ++n_mine) {
        const uint32_t g = g0 + (uint32_t)t;
        const uint32_t kst = g % kKsStages;
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        // warp-uniform: a warp with any row that needs the mask border takes the masked (per-element select) variant
        const bool need_mask =
            __any_sync(0xffffffffu, p.is_causal || !row_valid || (tile_start + kTileTokens > sg.valid_end));
        const bool tr_w = tr && quad == 0;
        if (gtid == 0) sm.trig[grp][(n_mine + 1) % 3] = 0;
        if (gtid == 0) B200_BEACON(5 + grp, g), B200_BEACON(8 + grp, 0), B200_BEACON(11, b);
        if (t <= 1) trace<TR>(tr_w, rank, 3, o_seg, 3);
        trace<TR>(tr_w, rank, 4, g, 3);

        // The leader's UMMA issuer saw ITS copy of the tile's k scales land before it issued this tile's Q.K^T, so in the
        // leader s_full vouches for them.  The peer CTA loads its own copy, and nothing there stands between that load
        // and the reads below: it waits on its own barrier (found by the timing-perturbation test: a late copy gave the
        // peer's 64 rows a tile computed with some of the slot's previous scales).  The slot is this group's own
        // (g + 6 = g mod 3), so the parity cannot alias.
        const float* ks_tile = s_ks + kst * 128;
        const float* lk_tile = sm.lkrk[grp];
        const float* rk_tile = sm.lkrk[grp] + 128;
        const uint32_t sb = g % kSBufs;
        const uint32_t s_addr = tmem + kColS + sb * 64 + lane_addr;
        uint8_t* prow = s_p + (2 * grp) * kPBytes + row * 128;  // this group's P pair: hi tile, lo tile right behind it
        bool fin_waited = false, fin_arrived = false;
        trace<TR>(tr_w, rank, 2, g, 0);
        mbar_wait(&sm.s_full[sb], (g / kSBufs) & 1);
        if (rank != 0) mbar_wait(&sm.ks_full[kst], (g / kKsStages) & 1);
        trace<TR>(tr_w, rank, 2, g, 1);
        tc_fence_after();

        // log2 and reciprocal of the tile's 128 k scales, once per tile for the whole group (a token past the end of the
        // sequence may hold anything: its reciprocal is forced to 0, its logit is masked by pass A)
        {
          const float k = ks_tile[gtid];
          const bool tok_ok = tile_start + gtid < sg.valid_end;
          sm.lkrk[grp][gtid] = __log2f(k);
          sm.lkrk[grp][128 + gtid] = tok_ok ? __frcp_rn(k) : 0.f;
        }
        named_bar_sync(gbar, 128);
        const float ymax = need_mask ? pass_max(std::true_type{}, s_addr, ks_tile, lk_tile, tile_start)
                                     : pass_max(std::false_type{}, s_addr, ks_tile, lk_tile, tile_start);
        const bool live = ymax > -INFINITY;
        trace<TR>(tr_w, rank, 4, g, 0);
        if (gtid == 0) B200_BEACON(8 + grp, 1);
        // the first tiles of a request always see a reference move (the first tile makes one): their weighted pass
        // waits for the predecessor's reference instead of running ahead of it (the max pass above did overlap it)
        // (the third tile only needs the first tile's reference, not its predecessor's publish: fin of tile g - 2 shares
        // this tile's barrier slot and stays complete until this group arrives on it; a move by tile g - 1 in between is
        // caught by the version check like anywhere else)
        if (t == 1 && g > 0) {
          mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);
          fin_waited = true;
        } else if (t == 2) {
          mbar_wait(&sm.fin[g & 1], ((g - 2) >> 1) & 1);
        }

        float acc = 0.f, w32 = 0.f;
        bool p_free_waited = false;
        // this group's P pair was last read by P.V of tile g - 3
        auto wait_p_free = [&]() {
          if (!p_free_waited) {
            mbar_wait(&sm.p_free[grp], ((g / kPPairs) & 1) ^ 1);
            p_free_waited = true;
          }
        };
        auto weighted_pass = [&]() {
          const float w = live ? ex2_approx(ymax - rloc) : 0.f;
          const uint32_t w_h2 = f32x2_to_f16x2(w, w);
          w32 = f16_lo_to_f32(w_h2);  // L carries the weight P.V sees
          wait_p_free();
          pass_codes(s_addr, rk_tile, ymax, w_h2, prow, acc);
        };
        // One code path for the pass over the codes (a single call site keeps the hot loop's footprint inside the
        // instruction cache): fast path = the current reference; if a row of the group has outgrown it (or on a
        // request's first tile) the reference moves first ("slow path") and the pass runs under the new one.
        bool trigger = (t == 0), moved = false;
        for (;;) {
          if (lane == 0) B200_BEACON(16 + (warp - 4), 11 + 100 * g);
          if (!trigger) {
            sync_view();
          } else {
            // ---- slow path: move the reference R of the rows that outgrew it, rescale O and the running sums ----
            if (!fin_waited && g > 0) mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);
            if (gtid == 0) B200_BEACON(8 + grp, 2);
            fin_waited = true;
            sync_view();
            // the two page maxima of a row meet in the group's own (free) P pair: it is written by the pass below
            wait_p_free();
            if (lane == 0) B200_BEACON(16 + (warp - 4), 21 + 100 * g);
            float* red1 = reinterpret_cast<float*>(s_p + (2 * grp) * kPBytes);  // [2 pages][64 rows]
            red1[page * 64 + row] = ymax;
            named_bar_sync(gbar, 128);
            // (== sm.st_ref[row] as of sync_view above: moves are serialised by fin.  Not read from shared memory here: the
            // row's other thread -- the lane of the second page, in another warp -- writes r_new below, and in the
            // deferred form nothing long stands between this line and that store: single rows came out wrong under the
            // timing-perturbation test until this read was gone)
            const float r_old = rloc;
            const float yb = fmaxf(red1[row], red1[64 + row]);
            const float r_new = yb > r_old ? yb + kRefHeadroom : r_old;
            const float fac = exp2f(r_old - r_new);  // 0 on the row's first move (R = -1e30), 1 if the row stays
            // Every P.V up to the previous tile must have landed in O before it is rescaled (P.V of tile g-1 is the last
            // one that can be in flight: tile g is not published yet and P.V is issued in tile order).  pv_done completes
            // once per tile; the parity wait could still be two phases behind if P.V of g-2 has not retired: wait for
            // that one first (s_free).
            auto drain_and_rescale = [&](float f_o) {
              if (lane == 0) B200_BEACON(16 + (warp - 4), 22 + 100 * g);
              if (g >= 2) mbar_wait(&sm.s_free[(g - 2) % kSBufs], ((g - 2) / kSBufs) & 1);
              if (lane == 0) B200_BEACON(16 + (warp - 4), 23 + 100 * g);
              mbar_wait(&sm.pv_done, (g - 1) & 1);
              tc_fence_after();
              if (__any_sync(0xffffffffu, f_o != 1.f)) rescale_o(tmem + kColO + lane_addr, f_o);
            };
            if (t != 0) {
              if constexpr (!DEFER) {
                drain_and_rescale(fac);
              } else {
                // DEFER: O is rescaled further down, behind this tile's weighted pass (which only needs the new
                // reference): the drain of the P.V products in flight -- a move's main cost -- then overlaps that pass.
                // The factor waits in shared memory, not in a register across the pass: the slot is where this thread's
                // row sum goes at the request's end; until then nobody reads it (the epilogue that read this buffer last,
                // two requests back, and a merge of the previous request are behind that request's barriers).
                sm.lred[rq & 1][slot][row] = fac;
              }
              lsum *= fac;
            }
            if (lane == 0) B200_BEACON(16 + (warp - 4), 24 + 100 * g);
            rloc = r_new;
            if (page == 0) sm.st_ref[row] = r_new;
            tc_fence_before();
            named_bar_sync(gbar, 128);  // O rescaled and the new row state written by the whole group
            ver += 1;
            if (gtid == 0) st_shared_volatile_u32(&sm.version, ver);  // reaches the other groups with the arrival on sm.fin
            // The reference is final for this tile and O is rescaled: the next tile may run its weighted pass while this
            // one runs its own (what orders them further -- P.V issue order, pv_done before a rescale -- does not hang
            // on fin).  The version store above is ordered before the arrival of gtid 0's warp by this barrier.
            // (DEFER: O is not rescaled yet.  Nothing of a later tile reaches O before this tile's P.V -- P.V is issued in
            // tile order and waits for this tile's publish, which follows the rescale below -- and a later tile that moves
            // the reference again waits for this tile's P.V, pv_done, before it touches O.)
            named_bar_sync(gbar, 128);
            if (lane == 0) B200_BEACON(16 + (warp - 4), 25 + 100 * g);
            if (lane == 0) mbar_arrive(&sm.fin[g & 1]);
            fin_arrived = true;
            moved = true;
          }
          if (lane == 0) B200_BEACON(16 + (warp - 4), 12 + 100 * g);
          const bool over = !moved && __any_sync(0xffffffffu, live && ymax > rloc);  // a row of this warp outgrew its reference
          if (!over) weighted_pass();
          if (lane == 0) B200_BEACON(16 + (warp - 4), 13 + 100 * g);
          if (moved) {
            if constexpr (DEFER) {
              if (t != 0) {
                const float f_o = sm.lred[rq & 1][slot][row];
                // (same waits, same order as the undeferred form: see drain_and_rescale in the slow path)
                if (g >= 2) mbar_wait(&sm.s_free[(g - 2) % kSBufs], ((g - 2) / kSBufs) & 1);
                mbar_wait(&sm.pv_done, (g - 1) & 1);
                tc_fence_after();
                if (__any_sync(0xffffffffu, f_o != 1.f)) rescale_o(tmem + kColO + lane_addr, f_o);
                tc_fence_before();  // (the publish below arrives on p_full behind this: P.V of this tile sees the rescaled O)
              }
            }
            lsum = fmaf(w32, acc, lsum);
            named_bar_sync(gbar, 128);  // lkrk is rewritten at this group's next tile
            break;
          }
          if (lane == 0) B200_BEACON(16 + (warp - 4), 14 + 100 * g);
          if (!fin_waited) {
            if (g > 0) mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);  // tile g-1 is final
            fin_waited = true;
          }
          trace<TR>(tr_w, rank, 4, g, 1);
          if (lane == 0) B200_BEACON(16 + (warp - 4), 15 + 100 * g);
          if (__shfl_sync(0xffffffffu, ld_shared_volatile_u32(&sm.version), 0) != ver) continue;  // computed under a reference that has moved
          if (lane == 0) B200_BEACON(16 + (warp - 4), 16 + 100 * g);
          if (over && lane == 0) sm.trig[grp][n_mine % 3] = 1;
          named_bar_sync(gbar, 128);
          trigger = sm.trig[grp][n_mine % 3] != 0;
          if (!trigger) {
            lsum = fmaf(w32, acc, lsum);
            break;
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.ks_empty[kst]);

        trace<TR>(tr_w, rank, 2, g, 2);
        if (gtid == 0) B200_BEACON(8 + grp, 3);
        // ---- publish P ----
        trace<TR>(tr_w, rank, 4, g, 2);
        // (token rows past the end of the sequence are zeroed in the V units by the warps that hand the units to P.V)
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(&sm.s_done[sb]);  // (the peer's share rides on its relay's p_relay hand-off)
          mbar_arrive(rank == 0 ? &sm.p_full[grp] : &sm.p_relay[grp]);
          if (!fin_arrived) mbar_arrive(&sm.fin[g & 1]);  // tile g is final: the next tile may publish
        }
        trace<TR>(tr_w, rank, 2, g, 3);
        if (gtid == 0) B200_BEACON(8 + grp, 4);
      }
      g0 += (uint32_t)sg.n_tiles;
      if (gtid == 0) B200_BEACON(8 + grp, 5), B200_BEACON(11, b);

      // every group has published its last tile: bring every thread's running sum to the request's final reference
      named_bar_sync(1, kSmThreads);
      {
        const float r_fin = ld_shared_volatile_f32(&sm.st_ref[row]);
        lsum *= exp2f(rloc - r_fin);
        rloc = r_fin;
      }
      // ---- epilogue of this request: O / L -> out (or split partials) ----
      // All three groups write the request out, each a third of the O columns (rounds of 32 columns taken round-robin):
      // a single group needs ~9k cycles for the 128 KB of an fp32 partial, during which its tiles of the next request
      // wait and the other groups run dry behind it (the tile hand-off is in order).
      sm.lred[rq & 1][slot][row] = lsum;
      // A request cut over several parts is merged by whichever of its parts reaches its epilogue last (no combine
      // launch): every part takes a ticket here; all but the last write their fp32 partial and count themselves done,
      // the last one waits for that count (its peers are already in their epilogues), merges in split order and
      // leaves both counters at zero for the next launch.
      int sp0 = 0, n_sp = 1;
      if (!sg.no_split) {
        sp0 = sg.split_idx - (b == begin_idx ? begin_n_split : 0);
        n_sp = (b < end_idx ? cur_ns : ns_end) - sp0;
      }
      // (small batches: requests in more than kMaxKernelMerge pieces are left to the combine kernel, see mla_common.cuh)
      const bool mrg = !sg.no_split && p.merge_state != nullptr && n_sp <= p.merge_limit;
      int* mstate = nullptr;
      if (mrg) {
        mstate = p.merge_state + ((((int64_t)b * gridDim.y + blockIdx.y) * 2 + rank) << 1);
        if (st_id == 0) sm.merge_ticket = atomicAdd(mstate, 1);
      }
      named_bar_sync(1, kSmThreads);
      const bool had_tiles = sg.n_tiles > 0;
      bool i_merge = false;
      if (mrg) {
        i_merge = sm.merge_ticket == n_sp - 1;
        if (i_merge) {
          // the other parts' partials were written long ago (a request's later parts run first) and have left L2 since:
          // start them on their way back now, while the last P.V of this part drains
          for (int i = 0; i < n_sp && i < 4; ++i) {
            if (sp0 + i == sg.split_idx) continue;
            if (st_id < n_valid_q) bulk_prefetch_l2(p.o_accum + ((int64_t)(sp0 + i) * p.M + q_row0 + st_id) * 512, 2048);
          }
        }
      }
      // the owner of the next request's first tile (global index g0 now) resets the row state; nobody reads the old one
      // any more, and the other groups look at it only after that tile is final (fin)
      if (b < end_idx && grp == (int)(g0 % kGroups) && page == 0) sm.st_ref[row] = kMInit;
      {
        const bool tr_e = tr && quad == 0 && grp == 0;
        trace<TR>(tr_e, rank, 3, o_seg, 0);
        const float* lr = &sm.lred[rq & 1][0][row];  // fixed summation order: results are bit-reproducible run to run
        const float l = ((lr[0] + lr[64]) + (lr[128] + lr[192])) + (lr[256] + lr[320]);
        const bool bad = (l == 0.f || l != l);
        const float inv = bad ? 1.f : 1.f / l;  // splitkv_mla.cu:1464-1466
        // merge: [row] the request's lse, [64 + row] this part's (log2), in the row-sum buffer of the other parity
        float* mg = &sm.lred[(rq & 1) ^ 1][0][0];
        if (i_merge) {
          if (st_id == 0) {  // the other parts of this request are in their epilogues already: a short wait
            B200_BEACON(12, 100 + n_sp);
            uint32_t spins = 0;
            while (ld_acquire_gpu_s32(mstate + 1) < n_sp - 1) {
              __nanosleep(64);
              if (++spins > (1u << 22)) __trap();  // a lost partial: fail loudly instead of hanging the device
            }
          }
          named_bar_sync(1, kSmThreads);
          if (grp == 0 && page == 0) {
            const float own = bad ? -INFINITY : log2f(l) + (rloc - kLog2Fp8Max);
            float glse = 0.f;
            if (row_valid) {  // the combine kernel's arithmetic (mla_combine.cu:72-110), splits in index order
              const float* la = p.lse_accum + (int64_t)sp0 * p.M + grow;
              const int own_i = sg.split_idx - sp0;
              float mx = -INFINITY;
              for (int i = 0; i < n_sp; ++i) mx = fmaxf(mx, i == own_i ? own : __ldcg(la + (int64_t)i * p.M));
              if (mx == -INFINITY) mx = 0.f;
              float sum = 0.f;
              for (int i = 0; i < n_sp; ++i) sum += exp2f((i == own_i ? own : __ldcg(la + (int64_t)i * p.M)) - mx);
              glse = (sum == 0.f || sum != sum) ? INFINITY : log2f(sum) + mx;
              const int sq = grow / p.h_q, hh = grow % p.h_q;
              p.lse[((int64_t)b * p.h_q + hh) * p.s_q + sq] = glse * 0.6931471805599453f;
            }
            mg[row] = glse;
            mg[64 + row] = own;
          }
          named_bar_sync(1, kSmThreads);
        }
        if (had_tiles) {  // (the merge's lse exchange above ran while the last P.V drained)
          if (gtid == 0) B200_BEACON(8 + grp, 7);
          // (not pv_done: it completes once per tile and a parity wait from here could be phases behind)
          mbar_wait(&sm.o_ready, o_seg & 1);
          trace<TR>(tr_e, rank, 3, o_seg, 1);
          tc_fence_after();
        }
        // O -> global through shared memory, so that every global store instruction writes whole 64 / 128-byte
        // segments (a TMEM lane is a q row: storing straight from the tcgen05.ld registers scatters 16-byte pieces
        // over 32 rows per instruction).  Staging = the group's own P pair (every P.V of this request has retired).
        // Warp-local: a warp's 32 lanes are 32 q rows of one TMEM lane half; it stages its [32 rows][32 columns] block
        // in a private 4 KB slice and reads it back row-major, so nothing but __syncwarp orders a round and the warps
        // hide each other's TMEM / shared / store latencies.
        uint8_t* stage_mine = s_p + (2 * grp + page) * kPBytes;
        const int64_t grow0 = (int64_t)b * p.M + q_row0;
        uint8_t* wst = stage_mine + (row >> 5) * 4096;
        const int wrow0 = row & 32;  // first of this warp's 32 rows within the CTA's 64
        auto flush = [&](const float* o, int r8, bool as_bf16, float inv) {
          const int d_base = 256 * (r8 >> 2) + 32 * (r8 & 3) + 128 * page;
          if (as_bf16) {
            uint8_t* srow = wst + lane * 64;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * k] * inv, o[8 * k + 1] * inv);
              __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * k + 2] * inv, o[8 * k + 3] * inv);
              __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * k + 4] * inv, o[8 * k + 5] * inv);
              __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * k + 6] * inv, o[8 * k + 7] * inv);
              *reinterpret_cast<uint4*>(srow + ((k ^ ((lane >> 1) & 3)) << 4)) =
                  make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                             *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
            }
            __syncwarp();
            const int k = lane & 3, rr = lane >> 2;  // per instruction: 8 rows x 64 B
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + rr;
              const uint4 v = *reinterpret_cast<const uint4*>(wst + r * 64 + ((k ^ ((r >> 1) & 3)) << 4));
              if (wrow0 + r < n_valid_q) *reinterpret_cast<uint4*>(p.out + (grow0 + wrow0 + r) * 512 + d_base + 8 * k) = v;
            }
          } else {
            uint8_t* srow = wst + lane * 128;
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<float4*>(srow + ((k ^ (lane & 7)) << 4)) =
                  make_float4(o[4 * k] * inv, o[4 * k + 1] * inv, o[4 * k + 2] * inv, o[4 * k + 3] * inv);
            __syncwarp();
            const int k = lane & 7, rr = lane >> 3;  // per instruction: 4 rows x 128 B
            if (!i_merge) {
              float* dst0 = p.o_accum + ((int64_t)sg.split_idx * p.M + q_row0 + wrow0) * 512 + d_base + 4 * k;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + rr;
                const float4 v = *reinterpret_cast<const float4*>(wst + r * 128 + ((k ^ (r & 7)) << 4));
                if (wrow0 + r < n_valid_q) *reinterpret_cast<float4*>(dst0 + (int64_t)r * 512) = v;
              }
            } else {
              // merge (mla_combine.cu:112-150): out = sum_i 2^(lse_i - lse) O_i over the request's splits in index order;
              // this part's O_i from the staging tile at its turn, the others' from their fp32 partials, 4 rows x 128 B
              // per load instruction and the 8 loads of a split in flight together
              float4 acc[8];
#pragma unroll
              for (int it = 0; it < 8; ++it) acc[it] = make_float4(0.f, 0.f, 0.f, 0.f);
              const int own_i = sg.split_idx - sp0;
#pragma unroll 1
              for (int i = 0; i < n_sp; ++i) {
                float4 x[8];
                float li[8];
                if (i == own_i) {
#pragma unroll
                  for (int it = 0; it < 8; ++it) {
                    const int r = it * 4 + rr;
                    x[it] = *reinterpret_cast<const float4*>(wst + r * 128 + ((k ^ (r & 7)) << 4));
                    li[it] = mg[64 + wrow0 + r];
                  }
                } else {
                  const int64_t prow0 = (int64_t)(sp0 + i) * p.M + q_row0 + wrow0;
#pragma unroll
                  for (int it = 0; it < 8; ++it) {
                    const int r = min(it * 4 + rr, n_valid_q - 1 - wrow0);  // rows past the request: a valid address, unused
                    x[it] = __ldcg(reinterpret_cast<const float4*>(p.o_accum + (prow0 + r) * 512 + d_base + 4 * k));
                    li[it] = __ldcg(p.lse_accum + prow0 + r);
                  }
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                  const float w = exp2f(li[it] - mg[wrow0 + it * 4 + rr]);
                  if (w == 0.f || w != w) continue;  // empty split: its partial may hold anything
                  acc[it].x = fmaf(w, x[it].x, acc[it].x), acc[it].y = fmaf(w, x[it].y, acc[it].y);
                  acc[it].z = fmaf(w, x[it].z, acc[it].z), acc[it].w = fmaf(w, x[it].w, acc[it].w);
                }
              }
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + rr;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(acc[it].x, acc[it].y), h1 = __floats2bfloat162_rn(acc[it].z, acc[it].w);
                if (wrow0 + r < n_valid_q)
                  *reinterpret_cast<uint2*>(p.out + (grow0 + wrow0 + r) * 512 + d_base + 4 * k) =
                      make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
              }
            }
          }
          __syncwarp();  // the slice is rewritten by the next round / this group's next P tile
        };
        float oa[32];
        const uint32_t o_addr = tmem + kColO + lane_addr;
#pragma unroll 1
        for (int r8 = grp; r8 < 8; r8 += kGroups) {
          if (had_tiles) {
            tmem_ld<32>(o_addr + r8 * 32, oa);
            tmem_wait_ld();
            if (r8 == grp) trace<TR>(tr_e, rank, 3, 8 + o_seg, 0);
            if (r8 + kGroups >= 8) {  // this warp's share of O has left TMEM: the next request's P.V may overwrite it
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(rank == 0 ? &sm.o_free : &sm.o_relay);
              trace<TR>(tr_e, rank, 3, 8 + o_seg, 2);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) oa[k] = 0.f;
          }
          flush(oa, r8, sg.no_split, inv);
          if (r8 == grp) trace<TR>(tr_e, rank, 3, 8 + o_seg, 1);
        }
        trace<TR>(tr_e, rank, 3, 8 + o_seg, 3);
        named_bar_sync(gbar, 128);  // the slices are this group's next P tiles (written per q row across warps)
        if (gtid == 0) B200_BEACON(8 + grp, 8);
        if (had_tiles) trace<TR>(tr_e, rank, 3, o_seg, 2);
        if (row_valid && page == 0 && grp == 0) {
          const float lse2 = log2f(l) + (rloc - kLog2Fp8Max);  // L_true = L * 2^R / 448
          if (sg.no_split) {
            const int sq = grow / p.h_q, hh = grow % p.h_q;
            p.lse[((int64_t)b * p.h_q + hh) * p.s_q + sq] = bad ? INFINITY : lse2 * 0.6931471805599453f;  // :1484
          } else if (!i_merge) {
            p.lse_accum[(int64_t)sg.split_idx * p.M + grow] = bad ? -INFINITY : lse2;                    // :1506
          }
        }
        if (mrg) {
          if (!i_merge) __threadfence();  // this thread's share of the partial, before the count below
          named_bar_sync(1, kSmThreads);  // (also: mg and merge_ticket are free for the next request)
          if (st_id == 0) {
            if (i_merge) {
              mstate[0] = 0, mstate[1] = 0;
            } else {
              atomicAdd(mstate + 1, 1);
            }
          }
        }
      }
      if (had_tiles) ++o_seg;
    }
  }

  if (TR && tid == 128 && part < 128 && blockIdx.y == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_part_cycles[(part * 2 + rank) * 4 + 1] = t;
    g_part_cycles[(part * 2 + rank) * 4 + 3] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA frees TMEM or exits while its peer can still address its barriers / shared memory
  if (warp == 1) tmem_dealloc_pair(tmem, kTmemCols);
}

}  // namespace pairk

int launch_mla_pair(const MlaParams& prm, int num_sm_parts, bool pdl, cudaStream_t stream) {
  const int smem = pairk::kTotal + 1024;
  const int q_pairs = (prm.M + 127) / 128;
  const dim3 grid(2, q_pairs, num_sm_parts), block(128 + pairk::kSmThreads);
  // B200_MLA_DEFER_RESCALE=1: the instantiation that rescales O behind a moving tile's weighted pass.  On logits whose
  // row maxima keep growing (spread of 4-6 bits: 14-40 % of the tiles move a reference) it is 10-19 % faster; on the
  // bench's logits (0.1 % of the tiles) it measures 2 % slower, so it is not the default.
  static const bool defer = getenv("B200_MLA_DEFER_RESCALE") && atoi(getenv("B200_MLA_DEFER_RESCALE")) != 0;
  auto launch = [&](auto* kernel) {
    B200_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    return launch_kernel(kernel, grid, block, smem, stream, pdl, prm);
  };
  if (prm.trace & 1) {  // diagnostic instantiation (B200_MLA_TRACE): pipeline timestamps, bit 2 = random delays
    static const int sites = getenv("B200_MLA_CHAOS_SITES") ? atoi(getenv("B200_MLA_CHAOS_SITES")) : 0x3ffffff;
    const int chaos = (prm.trace & 4) ? sites : 0;
    B200_CUDA(cudaMemcpyToSymbolAsync(pairk::g_chaos, &chaos, sizeof(int), 0, cudaMemcpyHostToDevice, stream));
    return defer ? launch(pairk::mla_decode_pair_kernel<true, true>) : launch(pairk::mla_decode_pair_kernel<true, false>);
  }
  return defer ? launch(pairk::mla_decode_pair_kernel<false, true>) : launch(pairk::mla_decode_pair_kernel<false, false>);
}

}  // namespace b200

// Diagnostic: copy the pipeline trace of the last traced pair-kernel launch (see g_trace) to host memory.
extern "C" int b200_mla_read_part_times(unsigned long long* out, int n) {
  const int total = 2 * 128 * 4;
  B200_CUDA(cudaMemcpyFromSymbol(out, b200::pairk::g_part_cycles, sizeof(unsigned long long) * (n < total ? n : total)));
  return 0;
}
extern "C" int b200_mla_read_trace(unsigned long long* out, int n) {
  const int total = 2 * b200::pairk::kTraceRoles * b200::pairk::kTraceTiles * 4;
  B200_CUDA(cudaMemcpyFromSymbol(out, b200::pairk::g_trace, sizeof(unsigned long long) * (n < total ? n : total)));
  return 0;
}

#ifdef B200_WAIT_DEBUG
#include <cstddef>
#include <cstdio>
// debug build only: where the first abandoned wait is recorded (pinned host memory, 8 unsigned)
extern "C" int b200_debug_set_record(unsigned* pinned) {
  cudaMemcpyToSymbol(b200::g_wait_host, &pinned, sizeof(pinned));
  unsigned z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaMemcpyToSymbol(b200::g_wait_dead, z, sizeof(z));
  using S = b200::pairk::Small;
#define OFF(f) printf("  %-10s %zu\n", #f, offsetof(S, f))
  OFF(k_full); OFF(k_empty); OFF(v_land); OFF(v_empty); OFF(ks_full); OFF(ks_empty); OFF(q_full); OFF(q_empty);
  OFF(s_full); OFF(s_free); OFF(s_done); OFF(v_peer); OFF(p_full); OFF(p_free); OFF(pv_done); OFF(o_ready); OFF(o_free);
  OFF(p_relay); OFF(o_relay); OFF(fin); OFF(tmem_base);
  printf("  Small at dynamic shared offset %d (+ alignment), CTA window 0x1000000 per rank\n", b200::pairk::kOffSmall);
  return 0;
}
#endif
