// Shared between the two MLA decode kernels (mla_decode.cu: SwapAB single-CTA, M <= 64 q rows; mla_decode_pair.cu:
// CTA-pair cta_group::2, 64 < M): launch parameters and the split-KV work partition
// (reference: ckv_fp8_per_token/splitkv_mla.cu:1218-1269, params.h).
#pragma once
#include <cuda_bf16.h>
#include <math.h>

#include "b200_decode.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace b200 {

static constexpr int kTileTokens = 128;
static constexpr int kPage = B200_MLA_PAGE_SIZE;
static constexpr float kMInit = -1e30f;                       // splitkv_mla.cu:16 (MAX_INIT_VAL_SM)
// Batches smaller than the number of scheduler parts (long sequences cut over dozens of parts): a request cut into
// more than this many pieces keeps the fp32 partials + combine path -- a single CTA merging them would serialise what
// the combine kernel spreads over the device (measured at b = 1, seq 32K: 79 us against 43.5 us).  Batches of at least
// one request per part (where a request is in two, rarely three pieces) merge everything in the kernel and launch no
// combine at all (an empty combine launch costs 3.5 us at the metric shape).
static constexpr int kMaxKernelMerge = 3;
static constexpr float kLog2Fp8Max = 8.807354922057604f;      // log2(448): a page's largest P' maps to the e4m3 maximum (splitkv_mla.cu:433)
// log2 of the growth a moved row reference leaves before the next move.  A move drains the P.V pipeline and rescales O
// in TMEM (2-3 tiles' worth of time), and with attention logits of spread sigma = 2.5 (log2 units: DeepSeek's 192^-0.5
// scaling on unit-variance q, k) a headroom of 4 sent 23 % of the tiles down that path, 7 sends 0.1 % (sigma = 4: 62 % /
// 14 %).  The price is e4m3 range: page weights down to 2^-7 leave the lo term of small codes subnormal; simulated
// distance to the faithful oracle 1.0-1.9e-3 instead of 0.7-1.0e-3 (tools/sim_hilo.py), 2.6e-3 at 8, 8e-3 at 10.
static constexpr float kRefHeadroom = 7.0f;

struct MlaParams {
  CUtensorMap tm_q_nope, tm_q_rope, tm_k_nope, tm_k_rope;
  const float* q_scale;
  const float* k_scale;
  const int32_t* block_table;
  int64_t block_table_stride;
  const int32_t* seqlens;
  const int32_t* meta;
  const int32_t* num_splits;
  __nv_bfloat16* out;
  float* lse;
  float* o_accum;
  float* lse_accum;
  int* merge_state;  // pair kernel: per (request, q block, CTA) {ticket, done} counters, zero between launches; null =
                     // split requests are left to the combine kernel
  int merge_limit;   // pair kernel: requests cut into more pieces than this are left to the combine kernel
  int batch, s_q, h_q, M;
  float scale_log2;
  int is_causal;
  const uint8_t* k_lora_base;  // raw pool pointers (bulk L2 prefetch)
  const uint8_t* k_rope_base;
  int pf_mode;   // pair kernel's L2 prefetch of upcoming K pages: 0 tensor-map prefetch, 1 none, 2 bulk prefetch (default)
  int pf_tiles;  // pair kernel: L2 prefetch distance of the K pages, in 128-token tiles (>= 2)
  int q_pf_tiles;  // pair kernel: the next request's Q is prefetched into L2 this many tiles before the current one ends
  int ks_ahead;  // pair kernel: k scales are loaded this many tiles ahead of the softmax (1 .. kKsAhead)
  int trace;  // diagnostic: pair kernel records clock64 timestamps of its pipeline events (B200_MLA_TRACE=1)
};

// The work of one (sm part, request) pair.
struct Segment {
  int batch_idx, seqlen, start_block, end_block, n_tiles, valid_end, split_idx;
  bool no_split;
};

__device__ __forceinline__ int mask_len(int s_q, int h_q, int M, int row) {  // splitkv_mla.cu:1136-1145
  return row < M ? s_q - row / h_q - 1 : 0;
}

// seqlen / num_splits_b: cache_seqlens[batch_idx] and num_splits[batch_idx], loaded by the caller (so that they can be
// fetched ahead of time: without an L1 every __ldg is an L2 round trip)
__device__ __forceinline__ Segment make_segment_from(const MlaParams& p, int batch_idx, int seqlen, int num_splits_b,
                                                     int begin_idx, int begin_seqlen, int end_idx, int end_seqlen,
                                                     int begin_n_split, int q_row0, int hb) {
  Segment s;
  s.batch_idx = batch_idx;
  s.seqlen = seqlen;
  s.start_block = batch_idx == begin_idx ? begin_seqlen / kPage : 0;
  const int full_blocks = (s.seqlen + kPage - 1) / kPage;
  int end_block = batch_idx == end_idx ? (end_seqlen + kPage - 1) / kPage : full_blocks;
  s.no_split = (s.start_block == 0 && end_block == full_blocks);  // splitkv_mla.cu:1239
  if (p.is_causal) {                                              // splitkv_mla.cu:1242-1266
    const int last_row = min(q_row0 + hb, p.M) - 1;
    int lim = s.seqlen - mask_len(p.s_q, p.h_q, p.M, last_row);
    if (batch_idx == end_idx) lim = min(lim, end_seqlen);
    end_block = (max(lim, 0) + kPage - 1) / kPage;
  }
  s.end_block = max(end_block, s.start_block);
  s.n_tiles = (s.end_block - s.start_block + 1) / 2;
  s.valid_end = min(s.seqlen, s.end_block * kPage);
  const int n_split_idx = batch_idx == begin_idx ? begin_n_split : 0;
  s.split_idx = num_splits_b + n_split_idx;  // splitkv_mla.cu:1490
  return s;
}

__device__ __forceinline__ Segment make_segment(const MlaParams& p, int batch_idx, int begin_idx, int begin_seqlen,
                                                int end_idx, int end_seqlen, int begin_n_split, int q_row0, int hb) {
  return make_segment_from(p, batch_idx, __ldg(p.seqlens + batch_idx), __ldg(p.num_splits + batch_idx), begin_idx,
                           begin_seqlen, end_idx, end_seqlen, begin_n_split, q_row0, hb);
}

// CTA-pair kernel for 64 < M (mla_decode_pair.cu); the Q tensor maps must use 64-row boxes.
int launch_mla_pair(const MlaParams& prm, int num_sm_parts, bool pdl, cudaStream_t stream);

}  // namespace b200

// Split-KV combine launch (small_kernels.cu); requests cut into at most `merged_up_to` pieces are skipped.
int launch_mla_combine(const float* o_accum, const float* lse_accum, const int32_t* num_splits, void* out, float* lse,
                       int batch, int s_q, int h_q, int merged_up_to, int use_pdl, void* stream);
