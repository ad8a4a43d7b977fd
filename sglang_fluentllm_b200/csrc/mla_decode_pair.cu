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
//   weight 2^(ymax - R) rides on the codes as an e4m3 hi + lo pair (two P tiles, two P.V products per tile).  Two
//   groups of four warps ping-pong over the tiles (group = tile parity); a tile's P depends on its predecessor only
//   through R, so the groups meet at a per-tile hand-off (fin) and a version counter; a group that computed under a
//   reference that has since moved recomputes its second pass.  Both passes over S are compact 16-column loops (TMEM
//   loads double-buffered, packed fp32x2 / f16x2 arithmetic, codes written straight into the swizzled P tiles).
//
//   Warp roles per CTA (384 threads): warp 0 = TMA producer for Q, K pages (-> leader barrier via cta_group::2 TMA)
//   and the k-scale ring; warp 2 = V producer; leader CTA: warp 1 issues every Q.K^T and warp 3 every P.V (two
//   independent instruction streams into the in-order tensor pipe; Q.K^T runs up to four tiles ahead), other CTA:
//   warp 1 relays the local P / O hand-offs to the leader; warp 1 also owns the TMEM allocation; warps 4-11 =
//   softmax / correction / epilogue.
//   TMEM per CTA: 4 S tiles (4 x 64 columns) + O (2 x 128 columns) = 512 columns.
//   Shared memory per CTA (226 KB): Q 40 KB, 2 K stages x 40 KB, 2 V stages x 32 KB, 2 P pairs (hi, lo) x 2 x 8 KB
//   (also the epilogue's staging tile), k-scale ring 4 KB, barriers / row state / log2 and reciprocal scales 6 KB.
#include <string.h>

#include "mla_common.cuh"

namespace b200 {
namespace pairk {

constexpr int kKStages = 2, kVStages = 2;
constexpr int kKsStages = 8, kKsAhead = 6;  // k-scale ring: 512 B per tile, loaded kKsAhead tiles ahead of the softmax
constexpr int kQBytes = 64 * 640;          // 4 nope chunks [64 rows x 128 B] + rope [64 x 128 B]
constexpr int kKStageBytes = 64 * 640;     // this CTA's page of the tile, same layout
constexpr int kVStageBytes = 2 * 128 * 128;  // [2 d-halves][128 tokens][128 B]
constexpr int kPBytes = 64 * 128;          // [64 q rows][128 tokens] e4m3, K-major SW128
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kQBytes;
constexpr int kOffV = kOffK + kKStages * kKStageBytes;
constexpr int kOffP = kOffV + kVStages * kVStageBytes;
constexpr int kPPairs = 2;                           // P tiles come in (hi, lo) pairs, one pair per softmax group (tile parity)
constexpr int kOffKs = kOffP + kPPairs * 2 * kPBytes;  // k-scale ring: [kKsStages][128] f32; the four P tiles (32 KB)
                                                     // double as the epilogue's staging tile
constexpr int kOffSmall = kOffKs + kKsStages * 512;
constexpr int kTotal = kOffSmall + 6144;
constexpr int kSmThreads = 256;                      // softmax group
constexpr int kSBufs = 4;                            // S tiles in TMEM: Q.K^T may run kSBufs - 1 tiles ahead of P.V
constexpr uint32_t kColS = 0, kColO = kSBufs * 64, kTmemCols = 512;
static_assert(kColO + 256 <= kTmemCols, "TMEM budget");
static_assert(kOffK % 1024 == 0 && kOffV % 1024 == 0 && kOffP % 1024 == 0 && kOffKs % 1024 == 0, "swizzle alignment");

struct Small {
  uint64_t k_full[kKStages], k_empty[kKStages];
  uint64_t v_land[kVStages], v_empty[kVStages];
  uint64_t ks_full[kKsStages], ks_empty[kKsStages];
  uint64_t q_full, q_empty;
  uint64_t s_full[kSBufs], s_free[kSBufs];     // s_free(g): P.V of tile g retired -> S tile g % kSBufs may be overwritten
  uint64_t p_full[kPPairs], p_free[kPPairs];   // per P pair (tile parity)
  uint64_t pv_done, o_ready, o_free;
  uint64_t p_relay[kPPairs], o_relay;  // non-leader CTA: local hand-off points forwarded to the leader by its warp 1
  uint64_t fin[2];               // tile g (buffer g & 1) is final: published, reference moves done
  uint32_t tmem_base;
  uint32_t version;              // bumped by every reference move
  uint32_t k_issued;             // tiles whose K page this CTA's producer has issued (orders the V re-read behind it)
  uint32_t trig[2][3];
  float st_ref[64];              // per-row reference exponent R (log2): O and L are in units of 2^R / 448
  float red1[2][2][64], lred[2][4][64];  // page maxima exchange; lred: row-sum exchange, double-buffered by request parity
  alignas(16) float lkrk[2][256];  // per softmax group: log2(k_scale) [128] and 1 / k_scale [128] of the tile it works on
};
static_assert(sizeof(Small) <= 6144, "Small too large");
static_assert(kTotal + 1024 <= 232448, "over the 227 KB shared-memory limit");

// Pipeline trace of part 0 (diagnostic): [rank][role: 0 producer, 1 mma, 2 softmax warp 4, 3 epilogue (indexed by
// request)][tile < 64][event < 4]
constexpr int kTraceTiles = 64;
constexpr int kTraceRoles = 5;
__device__ unsigned long long g_trace[2 * kTraceRoles * kTraceTiles * 4];
__device__ unsigned long long g_part_cycles[2 * 128 * 4];  // [part < 128][rank]{globaltimer ns at entry, at exit, clock64 at entry, at exit}
// TR = false (the production instantiation) compiles every trace site out
template <bool TR>
__device__ __forceinline__ void trace(bool on, uint32_t rank, int role, uint32_t g, int ev) {
  if constexpr (TR) {
    if (on && g < (uint32_t)kTraceTiles) g_trace[((rank * kTraceRoles + role) * kTraceTiles + g) * 4 + ev] = clock64();
  }
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


template <bool TR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
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

  const int32_t* meta = p.meta + part * B200_MLA_METADATA_INTS;
  const int begin_idx = __ldg(meta + 0), begin_seqlen = __ldg(meta + 1), end_idx = __ldg(meta + 2),
            end_seqlen = __ldg(meta + 3), begin_n_split = __ldg(meta + 4);
  // the split-KV combine kernel behind this one is launched with programmatic stream serialization: its CTAs may be
  // scheduled as soon as SMs drain (its griddepcontrol.wait orders the reads of the partials)
  pdl_launch_dependents();
  if (begin_idx >= p.batch) return;  // splitkv_mla.cu:1226 (both CTAs of the pair take the same exit)

  if (tid == 0) {
    for (int i = 0; i < kKStages; ++i) {
      mbar_init(&sm.k_full[i], 2);  // one expect_tx arrival per CTA of the pair (leader's copy is the live one)
      mbar_init(&sm.k_empty[i], 1);
    }
    for (int i = 0; i < kVStages; ++i) {
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
    }
    for (int i = 0; i < kPPairs; ++i) {
      mbar_init(&sm.p_full[i], 5);   // the four warps of the leader's softmax group + the peer's relay
      mbar_init(&sm.p_relay[i], 4);
      mbar_init(&sm.p_free[i], 1);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&sm.fin[i], 4);
    mbar_init(&sm.pv_done, 1);
    mbar_init(&sm.o_ready, 1);
    mbar_init(&sm.o_free, 5);   // the four warps of the leader's epilogue group + the peer's relay
    mbar_init(&sm.o_relay, 4);
    for (int i = 0; i < 3; ++i) sm.trig[0][i] = sm.trig[1][i] = 0;
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

  if (warp == 0) {
    // ============================== TMA producer (both CTAs) ==============================
    const uint32_t q_full_l = mapa_u32(smem_u32(&sm.q_full), 0);
    uint32_t k_full_l[kKStages];
#pragma unroll
    for (int i = 0; i < kKStages; ++i) k_full_l[i] = mapa_u32(smem_u32(&sm.k_full[i]), 0);

    uint32_t g = 0, seg_i = 0;
    const int pf = min(p.pf_tiles, 14);  // the block-table window holds 32 entries past the current tile
    const int ks_ahead = p.ks_ahead;
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
  } else if (TR && warp == 3 && rank == 1 && part == 0 && blockIdx.y == 0) {
    // diagnostic: when each V stage lands in the non-leader CTA
    for (uint32_t g = 0; g < 40; ++g) {
      mbar_wait(&sm.v_land[g % kVStages], (g / kVStages) & 1);
      trace<TR>(lane == 0, rank, 0, g, 3);
    }
  } else if (warp == 2) {
    // ============================== V producer (both CTAs) ==============================
    // Its own warp: the V stage of tile t frees when P.V of tile t-2 retires, which must not hold back the K stream
    // (Q.K^T runs ahead of P.V).
    const uint64_t pol_last_use = policy_evict_first();
    uint32_t g = 0;
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
      // both pages of tile t, this CTA's two 128-dim output slices (+ the tile's k scales) -> V stage (local barrier);
      // these bytes were just streamed through L2 as K by this CTA or its peer
      auto load_v = [&](int t, uint32_t gt) {
        const uint32_t st = gt % kVStages;
        trace<TR>(tr, rank, 0, gt, 1);
        mbar_wait(&sm.v_empty[st], ((gt / kVStages) & 1) ^ 1);
        // behind this tile's K issue: the same bytes then come out of L2 (or merge with the K fetch in flight) instead
        // of competing with the K and Q loads for HBM at a request's start
        for (uint32_t spins = 0; !(TR && (p.trace & 32)) && ld_shared_volatile_u32(&sm.k_issued) <= gt; ++spins) {
          if (spins > B200_WAIT_SPIN_LIMIT) __trap();
          __nanosleep(64);
        }
        trace<TR>(tr, rank, 0, gt, 2);
        const int blk0 = sg.start_block + 2 * t;
        const int npages = min(2, sg.end_block - blk0);
        const int tok0 = bw.get(blk0) * kPage;
        const int tok1 = npages > 1 ? bw.get(blk0 + 1) * kPage : 0;
        uint8_t* stage = s_v + st * kVStageBytes;
        if (TR && (p.trace & 8)) {  // diagnostic (wrong results): no V traffic at all
          if (elect_one()) mbar_arrive(&sm.v_land[st]);
          __syncwarp();
          return;
        }
        if (elect_one()) {
          mbar_arrive_expect_tx(&sm.v_land[st], (uint32_t)npages * (2 * 8192));
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_load_2d_hint(stage + j * 16384, &p.tm_k_nope, &sm.v_land[st], 256 * j + 128 * (int)rank, tok0, pol_last_use);
          if (npages > 1) {
#pragma unroll
            for (int j = 0; j < 2; ++j)
              tma_load_2d_hint(stage + j * 16384 + 8192, &p.tm_k_nope, &sm.v_land[st], 256 * j + 128 * (int)rank, tok1,
                               pol_last_use);
          }
        }
        __syncwarp();
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
    // issuer needs either a fixed interleave or polling, and both cost throughput).  Q.K^T of tile g waits for P.V of
    // tile g - kSBufs to retire (s_free): S tile g % kSBufs has then been read by its softmax group.
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
          if (g >= (uint32_t)kSBufs) mbar_wait(&sm.s_free[sb], ((g / kSBufs) & 1) ^ 1);  // P.V of tile g - kSBufs retired
          mbar_wait(&sm.ks_full[g % kKsStages], (g / kKsStages) & 1);  // s_full then vouches for the k scales too
          mbar_wait_cluster(&sm.k_full[kst], (g / kKStages) & 1);
          trace<TR>(tr, rank, 1, g, 0);
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
          const uint32_t pp = g & 1, vst = g % kVStages;
          mbar_wait_cluster(&sm.p_full[pp], (g >> 1) & 1);  // P written in both CTAs, V landed in the peer
          mbar_wait(&sm.v_land[vst], (g / kVStages) & 1);        // V landed here
          if (t == 0 && seg > 0) mbar_wait_cluster(&sm.o_free, (seg - 1) & 1);
          trace<TR>(tr, rank, 1, g, 2);
          tc_fence_after();
          const int tile_start = (sg.start_block + 2 * t) * kPage;
          const int nk = (min(kTileTokens, sg.valid_end - tile_start) + 31) / 32;
          if (elect_one()) {
#pragma unroll
            for (int hl = 0; hl < 2; ++hl)  // P = P_hi + P_lo
#pragma unroll
              for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int s = 0; s < 4; ++s)
                  if (s < nk)
                    umma_f8_pair(tmem + kColO + j * 128, with_addr(d_kmaj, p_addr + (2 * pp + hl) * kPBytes + s * 32),
                                 with_addr(d_vmn, v_addr + vst * kVStageBytes + j * 16384 + s * 4096), idesc_pv,
                                 !(t == 0 && s == 0 && hl == 0));
            umma_commit_pair(&sm.v_empty[vst], 3);
            umma_commit_pair(&sm.p_free[pp], 3);
            umma_commit_pair(&sm.s_free[g % kSBufs], 3);
            umma_commit_pair(&sm.pv_done, 3);
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
          mbar_wait(&sm.p_relay[g & 1], (g >> 1) & 1);
          mbar_wait(&sm.v_land[g % kVStages], (g / kVStages) & 1);  // the forwarded arrival also vouches for this CTA's V
          if (elect_one()) mbar_arrive_cluster(p_full_l[g & 1]);
          __syncwarp();
        }
        mbar_wait(&sm.o_relay, seg & 1);
        if (elect_one()) mbar_arrive_cluster(o_free_l);
        __syncwarp();
        ++seg;
      }
    }
  } else if (warp >= 4) {
    // ============================== softmax / correction / epilogue (both CTAs) ==============================
    // Two groups of four warps ping-pong over the tiles: group 0 (warps 4-7) serves the tiles whose S / P buffer is 0,
    // group 1 (warps 8-11) buffer 1.  A group's per-tile chain (barrier waits, TMEM load, vote, publish) then overlaps
    // the other group's exponentials instead of idling the SM.  Under the lazy reference a tile's P does not depend
    // on its predecessor unless a reference move ("trigger") happens, so the groups only meet at a per-tile hand-off
    // (sm.fin) that orders the moves: tile g is published after tile g-1 is final, and a group that computed under a
    // reference which has since moved (sm.version) recomputes.
    const int st_id = tid - 128;              // 0 .. 255 over both groups
    const int grp = (warp - 4) >> 2;          // S / P buffer (tile parity) this warp serves
    const int gtid = st_id & 127;             // thread within the group
    const uint32_t gbar = 2 + grp;            // the group's named barrier (128 threads)
    const int quad = warp & 3;                // TMEM lane quadrant of this warp
    const int tl = quad * 32 + lane;          // TMEM lane
    const int row = tl & 63;                  // q row within this CTA
    const int page = tl >> 6;                 // lanes 64-127 hold the second half of the N columns = the peer's page
    const int ch = grp;                       // epilogue only: which 64 of the 128 O columns of each d half
    const int slot = page * 2 + ch;           // epilogue only: which quarter of the staged row this thread fills
    const uint32_t lane_addr = (uint32_t)(quad * 32) << 16;
    const bool row_valid = row < n_valid_q;
    const int grow = q_row0 + row;            // row within the request's M rows
    uint32_t g0 = 0, o_seg = 0;               // g0: global index of the request's first tile
    uint32_t n_mine = 0;                      // tiles this group has served (trigger-flag ring)
    // per-request scalars, fetched one request ahead (three dependent L2 round trips otherwise sit between a request's
    // epilogue and the next request's first tile)
    int cur_seqlen = __ldg(p.seqlens + begin_idx), cur_ns = __ldg(p.num_splits + begin_idx);
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
        const uint32_t v_now = ld_shared_volatile_u32(&sm.version);
        if (v_now != ver) {
          const float r_new = ld_shared_volatile_f32(&sm.st_ref[row]);
          lsum *= exp2f(rloc - r_new);
          rloc = r_new;
          ver = v_now;
        }
      };

      // Pass A over this lane's 64 S columns: ymax = max_j y_j, y_j = s_j f k_j + log2 k_j = log2 of the un-normalised
      // P'_j (splitkv_mla.cu:400-432 computes the same maximum as max_j exp2(..) k_j).  16 columns at a time, TMEM loads
      // double-buffered, a compact loop (the kernel's hot code stays inside the instruction cache).
      auto pass_max = [&](uint32_t s_addr, const float* ks_tile, const float* lk_tile, int tile_start, bool need_mask) {
        const float2 f2 = make_float2(f, f);
        const float4* ksrc = reinterpret_cast<const float4*>(ks_tile + page * 64);
        const float4* lsrc = reinterpret_cast<const float4*>(lk_tile + page * 64);
        const int nvalid = lim - (tile_start + page * 64);  // columns of this lane below the mask border
        float m0 = -INFINITY, m1 = -INFINITY;
        auto chunk = [&](const float* x, int c16) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 k4 = ksrc[c16 * 4 + i], l4 = lsrc[c16 * 4 + i];
            if (!need_mask) {
              const float2 ya = fma2(make_float2(x[4 * i], x[4 * i + 1]), mul2(f2, make_float2(k4.x, k4.y)), make_float2(l4.x, l4.y));
              const float2 yb = fma2(make_float2(x[4 * i + 2], x[4 * i + 3]), mul2(f2, make_float2(k4.z, k4.w)), make_float2(l4.z, l4.w));
              m0 = fmaxf(m0, fmaxf(ya.x, ya.y));
              m1 = fmaxf(m1, fmaxf(yb.x, yb.y));
            } else {
              const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float y = fmaf(x[4 * i + j], f * kk[j], ll[j]);
                if (c16 * 16 + 4 * i + j < nvalid) m0 = fmaxf(m0, y);  // a select: poisoned padding never propagates
              }
            }
          }
        };
        float xa[16], xb[16];
        tmem_ld16(s_addr, reinterpret_cast<uint32_t*>(xa));
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16 += 2) {
          tmem_wait_ld();
          tmem_ld16(s_addr + (c16 + 1) * 16, reinterpret_cast<uint32_t*>(xb));
          chunk(xa, c16);
          tmem_wait_ld();
          if (c16 == 0) tmem_ld16(s_addr + 32, reinterpret_cast<uint32_t*>(xa));
          chunk(xb, c16 + 1);
        }
        return fmaxf(m0, m1);
      };
      // Pass B: the page's codes code_j = e4m3(448 * 2^(y_j - ymax)) (:433-440), the page sum sum_j 448 * 2^(y_j - ymax)
      // / k_j (:426,434), and code_j * w as an e4m3 pair hi + lo straight into the two P tiles (K-major, 128-byte
      // swizzle; nobody reads them before this group arrives on p_full).
      auto pass_codes = [&](uint32_t s_addr, const float* ks_tile, const float* lk_tile, const float* rk_tile, int tile_start,
                            bool need_mask, float ymax, uint32_t w_h2, uint8_t* prow, float& acc) {
        const float2 f2 = make_float2(f, f);
        const float c = (ymax > -INFINITY ? ymax : 0.f) - kLog2Fp8Max;
        const float2 nc2 = make_float2(-c, -c), one2 = make_float2(1.f, 1.f);
        float2 acc2 = make_float2(0.f, 0.f);
        const float4* ksrc = reinterpret_cast<const float4*>(ks_tile + page * 64);
        const float4* lsrc = reinterpret_cast<const float4*>(lk_tile + page * 64);
        const float4* rsrc = reinterpret_cast<const float4*>(rk_tile + page * 64);
        const int nvalid = lim - (tile_start + page * 64);
        auto chunk = [&](const float* x, int c16) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 k4 = ksrc[c16 * 4 + i], l4 = lsrc[c16 * 4 + i], r4 = rsrc[c16 * 4 + i];
            float v[4];
            if (!need_mask) {
              const float2 ya = fma2(make_float2(x[4 * i], x[4 * i + 1]), mul2(f2, make_float2(k4.x, k4.y)), make_float2(l4.x, l4.y));
              const float2 yb = fma2(make_float2(x[4 * i + 2], x[4 * i + 3]), mul2(f2, make_float2(k4.z, k4.w)), make_float2(l4.z, l4.w));
              const float2 aa = fma2(ya, one2, nc2), ab = fma2(yb, one2, nc2);
              v[0] = ex2_approx(aa.x), v[1] = ex2_approx(aa.y), v[2] = ex2_approx(ab.x), v[3] = ex2_approx(ab.y);
              acc2 = fma2(make_float2(v[0], v[1]), make_float2(r4.x, r4.y), acc2);
              acc2 = fma2(make_float2(v[2], v[3]), make_float2(r4.z, r4.w), acc2);
            } else {
              const float kk[4] = {k4.x, k4.y, k4.z, k4.w}, ll[4] = {l4.x, l4.y, l4.z, l4.w}, rr[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool ok = c16 * 16 + 4 * i + j < nvalid;
                const float y = fmaf(x[4 * i + j], f * kk[j], ll[j]);
                v[j] = ok ? ex2_approx(y - c) : 0.f;
                acc2.x += ok ? v[j] * rr[j] : 0.f;
              }
            }
            uint32_t h0, l0, h1, l1;
            split_hi_lo(cvt_e4m3x2(v[0], v[1]), w_h2, h0, l0);
            split_hi_lo(cvt_e4m3x2(v[2], v[3]), w_h2, h1, l1);
            hi[i] = h0 | (h1 << 16);
            lo[i] = l0 | (l1 << 16);
          }
          const int off = ((page * 4 + c16) ^ (row & 7)) << 4;
          *reinterpret_cast<uint4*>(prow + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<uint4*>(prow + kPBytes + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        };
        // two register sets: the TMEM load of the next 16 columns is in flight while this one is processed
        float xa[16], xb[16];
        tmem_ld16(s_addr, reinterpret_cast<uint32_t*>(xa));
#pragma unroll 1
        for (int c16 = 0; c16 < 4; c16 += 2) {
          tmem_wait_ld();
          tmem_ld16(s_addr + (c16 + 1) * 16, reinterpret_cast<uint32_t*>(xb));
          chunk(xa, c16);
          tmem_wait_ld();
          if (c16 == 0) tmem_ld16(s_addr + 32, reinterpret_cast<uint32_t*>(xa));
          chunk(xb, c16 + 1);
        }
        acc = acc2.x + acc2.y;
      };

      const int t_first = (int)((grp ^ g0) & 1u);  // this group's first tile of the request
      for (int t = t_first; t < sg.n_tiles; t += 2, ++n_mine) {
        const uint32_t g = g0 + (uint32_t)t;
        const uint32_t vst = g % kVStages, kst = g % kKsStages;
        const int tile_start = (sg.start_block + 2 * t) * kPage;
        const bool need_mask = p.is_causal || !row_valid || (tile_start + kTileTokens > sg.valid_end);
        const bool tr_w = tr && quad == 0;
        if (gtid == 0) sm.trig[grp][(n_mine + 1) % 3] = 0;
        if (t <= 1) trace<TR>(tr_w, rank, 3, o_seg, 3);
        trace<TR>(tr_w, rank, 4, g, 3);

        // No wait on ks_full: the UMMA issuer saw the tile's k scales land before it issued this tile's Q.K^T, so
        // s_full vouches for them too.
        const float* ks_tile = s_ks + kst * 128;
        const float* lk_tile = sm.lkrk[grp];
        const float* rk_tile = sm.lkrk[grp] + 128;
        const uint32_t sb = g % kSBufs;
        const uint32_t s_addr = tmem + kColS + sb * 64 + lane_addr;
        uint8_t* prow = s_p + (2 * grp) * kPBytes + row * 128;  // this group's P pair: hi tile, lo tile right behind it
        // the first two tiles of a request always see a reference move: wait for the predecessor before computing
        bool fin_waited = false;
        if (t <= 1 && g > 0) {
          mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);
          fin_waited = true;
        }
        trace<TR>(tr_w, rank, 2, g, 0);
        mbar_wait(&sm.s_full[sb], (g / kSBufs) & 1);
        trace<TR>(tr_w, rank, 2, g, 1);
        tc_fence_after();

        // log2 and reciprocal of the tile's 128 k scales, once per tile for the whole group
        {
          const float k = ks_tile[gtid];
          sm.lkrk[grp][gtid] = log2f(k);
          sm.lkrk[grp][128 + gtid] = 1.f / k;
        }
        named_bar_sync(gbar, 128);
        const float ymax = pass_max(s_addr, ks_tile, lk_tile, tile_start, need_mask);
        const bool live = ymax > -INFINITY;
        trace<TR>(tr_w, rank, 4, g, 0);

        float acc = 0.f, w32 = 0.f;
        bool p_free_waited = false;
        // this group's P pair was last read by P.V of tile g - 2
        auto wait_p_free = [&]() {
          if (!p_free_waited) {
            mbar_wait(&sm.p_free[grp], ((g >> 1) & 1) ^ 1);
            p_free_waited = true;
          }
        };
        auto weighted_pass = [&]() {
          const float w = live ? ex2_approx(ymax - rloc) : 0.f;
          const uint32_t w_h2 = f32x2_to_f16x2(w, w);
          w32 = f16_lo_to_f32(w_h2);  // L carries the weight P.V sees
          wait_p_free();
          pass_codes(s_addr, ks_tile, lk_tile, rk_tile, tile_start, need_mask, ymax, w_h2, prow, acc);
        };
        bool trigger = (t == 0);
        while (!trigger) {
          // ---- fast path: the current reference ----
          sync_view();
          const bool over = __any_sync(0xffffffffu, live && ymax > rloc);  // a row of this warp outgrew its reference
          if (!over) weighted_pass();
          if (!fin_waited) {
            if (g > 0) mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);  // tile g-1 is final
            fin_waited = true;
          }
          trace<TR>(tr_w, rank, 4, g, 1);
          if (ld_shared_volatile_u32(&sm.version) != ver) continue;  // computed under a reference that has moved
          if (over && lane == 0) sm.trig[grp][n_mine % 3] = 1;
          named_bar_sync(gbar, 128);
          trigger = sm.trig[grp][n_mine % 3] != 0;
          break;
        }
        if (!trigger) {
          lsum = fmaf(w32, acc, lsum);
        } else {
          // ---- slow path: move the reference R of the rows that outgrew it, rescale O and the running sums ----
          if (!fin_waited && g > 0) mbar_wait(&sm.fin[(g - 1) & 1], ((g - 1) >> 1) & 1);
          sync_view();
          sm.red1[grp][page][row] = ymax;
          named_bar_sync(gbar, 128);
          const float r_old = ld_shared_volatile_f32(&sm.st_ref[row]);  // == rloc
          const float yb = fmaxf(sm.red1[grp][0][row], sm.red1[grp][1][row]);
          const float r_new = yb > r_old ? yb + kRefHeadroom : r_old;
          const float fac = exp2f(r_old - r_new);  // 0 on the row's first move (R = -1e30), 1 if the row stays
          if (t != 0) {
            // every P.V up to the previous tile must have landed in O before it is rescaled (P.V of tile g-1 is the
            // last one issued: tile g is not published yet).  pv_done completes once per tile and Q.K^T runs ahead of
            // P.V: first make sure P.V of tile g-2 has retired (its p_free, at most one phase behind), so that the
            // parity wait on pv_done cannot alias the phase before it.
            wait_p_free();
            mbar_wait(&sm.pv_done, (g - 1) & 1);
            tc_fence_after();
            if (__any_sync(0xffffffffu, fac != 1.f)) {
#pragma unroll 1
              for (int c8 = 0; c8 < 8; ++c8) {
                float o[32];
                const uint32_t a = tmem + kColO + c8 * 32 + lane_addr;
                tmem_ld<32>(a, o);
                tmem_wait_ld();
#pragma unroll
                for (int k = 0; k < 32; ++k) o[k] *= fac;
                tmem_st<32>(a, o);
              }
              tmem_wait_st();
            }
            lsum *= fac;
          }
          rloc = r_new;
          if (page == 0) sm.st_ref[row] = r_new;
          tc_fence_before();
          named_bar_sync(gbar, 128);  // O rescaled and the new row state written by the whole group
          ver += 1;
          if (gtid == 0) st_shared_volatile_u32(&sm.version, ver);  // reaches the other group with this warp's arrival on sm.fin
          weighted_pass();
          lsum = fmaf(w32, acc, lsum);
          named_bar_sync(gbar, 128);  // lkrk is rewritten at this group's next tile
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.ks_empty[kst]);

        trace<TR>(tr_w, rank, 2, g, 2);
        // ---- publish P ----
        trace<TR>(tr_w, rank, 4, g, 2);
        // token rows past the end of the sequence meet P = 0 in P.V; make V there finite (0 * NaN).
        // Reference: fill_oob_V, splitkv_mla.cu:698-721.
        if (tile_start + kTileTokens > sg.valid_end) {
          mbar_wait(&sm.v_land[vst], (g / kVStages) & 1);
          const int valid_in_tile = sg.valid_end - tile_start;
          const int nk = (min(kTileTokens, valid_in_tile) + 31) / 32;
          if (gtid >= valid_in_tile && gtid < nk * 32) {
            uint8_t* stage = s_v + vst * kVStageBytes;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint4* vr = reinterpret_cast<uint4*>(stage + j * 16384 + gtid * 128);
#pragma unroll
              for (int k = 0; k < 8; ++k) vr[k] = make_uint4(0, 0, 0, 0);
            }
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(rank == 0 ? &sm.p_full[grp] : &sm.p_relay[grp]);
          mbar_arrive(&sm.fin[g & 1]);  // tile g is final: the other group may publish tile g+1
        }
        trace<TR>(tr_w, rank, 2, g, 3);
      }
      g0 += (uint32_t)sg.n_tiles;

      // both groups have published their last tile: bring every thread's running sum to the request's final reference
      named_bar_sync(1, kSmThreads);
      {
        const float r_fin = ld_shared_volatile_f32(&sm.st_ref[row]);
        lsum *= exp2f(rloc - r_fin);
        rloc = r_fin;
      }
      // ---- epilogue of this request: O / L -> out (or split partials) ----
      // One group writes the request out while the other starts the next request (its first tile always moves the
      // reference, ~3k cycles, and the fp32 partial stores drain at ~28 B/clk per SM, ~4.6k cycles): the epilogue group
      // is the one that does NOT own the next request's first tile (global index g0 now).
      const int egrp = b < end_idx ? (int)((g0 & 1u) ^ 1u) : 1;
      sm.lred[rq & 1][slot][row] = lsum;
      named_bar_sync(1, kSmThreads);
      const bool had_tiles = sg.n_tiles > 0;
      if (grp != egrp) {
        if (b < end_idx) {  // reset the row state for the next request; nobody reads the old one any more
          if (page == 0) sm.st_ref[row] = kMInit;
          named_bar_sync(gbar, 128);
        }
      } else {
        const bool tr_e = tr && quad == 0;
        trace<TR>(tr_e, rank, 3, o_seg, 0);
        const float l = (sm.lred[rq & 1][0][row] + sm.lred[rq & 1][1][row]) + (sm.lred[rq & 1][2][row] + sm.lred[rq & 1][3][row]);
        const bool bad = (l == 0.f || l != l);
        const float inv = bad ? 1.f : 1.f / l;  // splitkv_mla.cu:1464-1466
        if (had_tiles) {
          // (not pv_done: it completes once per tile and a parity wait from here could be phases behind)
          mbar_wait(&sm.o_ready, o_seg & 1);
          trace<TR>(tr_e, rank, 3, o_seg, 1);
          tc_fence_after();
        }
        // O -> global through shared memory, so that every global store instruction writes whole 64 / 128-byte
        // segments (a TMEM lane is a q row: storing straight from the tcgen05.ld registers scatters 16-byte pieces
        // over 32 rows per instruction).  Staging = this group's own P pair (its next tiles' buffers: every P.V of
        // this request has retired, and the other group only writes its own pair): one 8 KB half per TMEM lane half
        // (= 128-column block of d), [64 rows][32 values], 16-byte chunks XOR-swizzled; eight rounds of 32 columns.
        uint8_t* stage_mine = s_p + (2 * egrp + page) * kPBytes;
        const int64_t grow0 = (int64_t)b * p.M + q_row0;
        const int gw = warp & 3;  // warp within the group
        if (p.epi_mode) {
          // Warp-local staging: a warp's 32 lanes are 32 q rows of one TMEM lane half; it stages its [32 rows][32
          // columns] block in a private 4 KB slice of the group's P tiles and reads it back row-major, so every global
          // store instruction still writes whole 64 / 128-byte row segments but nothing but __syncwarp orders the
          // round: the four warps run independently and hide each other's TMEM / shared / store latencies (the
          // group-wide staging below spends ~900 cycles per round on two 128-thread barriers).
          uint8_t* wst = stage_mine + (row >> 5) * 4096;
          const int wrow0 = row & 32;  // first of this warp's 32 rows within the CTA's 64
          auto flush = [&](const float* o, int r8) {
            const int d_base = 256 * (r8 >> 2) + 32 * (r8 & 3) + 128 * page;
            if (sg.no_split) {
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
              float* dst0 = p.o_accum + ((int64_t)sg.split_idx * p.M + q_row0 + wrow0) * 512 + d_base + 4 * k;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = it * 4 + rr;
                const float4 v = *reinterpret_cast<const float4*>(wst + r * 128 + ((k ^ (r & 7)) << 4));
                if (wrow0 + r < n_valid_q) *reinterpret_cast<float4*>(dst0 + (int64_t)r * 512) = v;
              }
            }
            __syncwarp();  // the slice is rewritten by the next round / this group's next P tile
          };
          float oa[32], ob[32];
          if (had_tiles) {
            const uint32_t o_addr = tmem + kColO + lane_addr;
            tmem_ld<32>(o_addr, oa);
#pragma unroll 1
            for (int r8 = 0; r8 < 8; r8 += 2) {
              tmem_wait_ld();
              tmem_ld<32>(o_addr + (r8 + 1) * 32, ob);
              flush(oa, r8);
              tmem_wait_ld();
              if (r8 + 2 < 8) {
                tmem_ld<32>(o_addr + (r8 + 2) * 32, oa);
              } else {  // O has left TMEM: the next request's P.V may overwrite it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(rank == 0 ? &sm.o_free : &sm.o_relay);
              }
              flush(ob, r8 + 1);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) oa[k] = 0.f;
#pragma unroll 1
            for (int r8 = 0; r8 < 8; ++r8) flush(oa, r8);
          }
          named_bar_sync(gbar, 128);  // the slices are this group's next P tiles (written per q row across warps)
        } else {
  #pragma unroll 1
          for (int r8 = 0; r8 < 8; ++r8) {
            const int d_base = 256 * (r8 >> 2) + 32 * (r8 & 3);  // + 128 * page
            float o[32];
            if (had_tiles) {
              tmem_ld<32>(tmem + kColO + r8 * 32 + lane_addr, o);
              tmem_wait_ld();
              if (r8 == 7) {  // O is in registers / shared memory: the next request's P.V may overwrite it
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(rank == 0 ? &sm.o_free : &sm.o_relay);
              }
            } else {
  #pragma unroll
              for (int k = 0; k < 32; ++k) o[k] = 0.f;
            }
            if (sg.no_split) {
              uint8_t* srow = stage_mine + row * 64;
  #pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 h0 = __floats2bfloat162_rn(o[8 * k] * inv, o[8 * k + 1] * inv);
                __nv_bfloat162 h1 = __floats2bfloat162_rn(o[8 * k + 2] * inv, o[8 * k + 3] * inv);
                __nv_bfloat162 h2 = __floats2bfloat162_rn(o[8 * k + 4] * inv, o[8 * k + 5] * inv);
                __nv_bfloat162 h3 = __floats2bfloat162_rn(o[8 * k + 6] * inv, o[8 * k + 7] * inv);
                *reinterpret_cast<uint4*>(srow + ((k ^ ((row >> 1) & 3)) << 4)) =
                    make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                               *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
              }
            } else {
              uint8_t* srow = stage_mine + row * 128;
  #pragma unroll
              for (int k = 0; k < 8; ++k)
                *reinterpret_cast<float4*>(srow + ((k ^ (row & 7)) << 4)) =
                    make_float4(o[4 * k] * inv, o[4 * k + 1] * inv, o[4 * k + 2] * inv, o[4 * k + 3] * inv);
            }
            named_bar_sync(gbar, 128);
            if (sg.no_split) {
              // per instruction: 4 rows x 2 lane halves x 64 B
              const int k = lane & 3, pg = (lane >> 2) & 1, rr = lane >> 3;
              const uint8_t* src = s_p + (2 * egrp + pg) * kPBytes;
  #pragma unroll
              for (int it = 0; it < 4; ++it) {
                const int r = gw * 16 + it * 4 + rr;
                const uint4 v = *reinterpret_cast<const uint4*>(src + r * 64 + ((k ^ ((r >> 1) & 3)) << 4));
                if (r < n_valid_q) *reinterpret_cast<uint4*>(p.out + (grow0 + r) * 512 + d_base + 128 * pg + 8 * k) = v;
              }
            } else {
              // per instruction: 2 rows x 2 lane halves x 128 B
              const int k = lane & 7, pg = (lane >> 3) & 1, rr = lane >> 4;
              const uint8_t* src = s_p + (2 * egrp + pg) * kPBytes;
              float* dst0 = p.o_accum + ((int64_t)sg.split_idx * p.M + q_row0) * 512;
  #pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int r = gw * 16 + it * 2 + rr;
                const float4 v = *reinterpret_cast<const float4*>(src + r * 128 + ((k ^ (r & 7)) << 4));
                if (r < n_valid_q) *reinterpret_cast<float4*>(dst0 + (int64_t)r * 512 + d_base + 128 * pg + 4 * k) = v;
              }
            }
            named_bar_sync(gbar, 128);  // the staging halves are rewritten by the next round / this group's next P tile
          }
        }
        if (had_tiles) trace<TR>(tr_e, rank, 3, o_seg, 2);
        if (row_valid && page == 0) {
          const float lse2 = log2f(l) + (rloc - kLog2Fp8Max);  // L_true = L * 2^R / 448
          if (sg.no_split) {
            const int sq = grow / p.h_q, hh = grow % p.h_q;
            p.lse[((int64_t)b * p.h_q + hh) * p.s_q + sq] = bad ? INFINITY : lse2 * 0.6931471805599453f;  // :1484
          } else {
            p.lse_accum[(int64_t)sg.split_idx * p.M + grow] = bad ? -INFINITY : lse2;                    // :1506
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

int launch_mla_pair(const MlaParams& prm, int num_sm_parts, cudaStream_t stream) {
  const int smem = pairk::kTotal + 1024;
  const int q_pairs = (prm.M + 127) / 128;
  if (prm.trace & 1) {  // diagnostic instantiation (B200_MLA_TRACE): pipeline timestamps, knobs that break results
    B200_CUDA(cudaFuncSetAttribute(pairk::mla_decode_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pairk::mla_decode_pair_kernel<true><<<dim3(2, q_pairs, num_sm_parts), 384, smem, stream>>>(prm);
  } else {
    B200_CUDA(cudaFuncSetAttribute(pairk::mla_decode_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    pairk::mla_decode_pair_kernel<false><<<dim3(2, q_pairs, num_sm_parts), 384, smem, stream>>>(prm);
  }
  B200_CUDA(cudaGetLastError());
  return 0;
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
