// The HBM-bound helper kernels around the MLA decode / FP8 GEMM hot path:
//   * tile-scheduler metadata (integer, bit-exact vs oracle/mla.py:get_mla_metadata)
//   * per-(token,head) FP8 quantisers for q and for new K tokens (scatter into the paged pool)
//   * indexed de-quantisation of the pool
//   * split-KV combine
//   * 1x128 activation quantiser for the block-scaled GEMM
// All are byte/elementwise streaming kernels: 128-bit coalesced loads/stores, one warp per row, no shared-memory
// staging needed (no reuse).  Arithmetic is IEEE fp32 (explicit _rn intrinsics, no FMA contraction) so e4m3 codes
// and scale bits are bit-exact against the numpy oracle.
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <math.h>

#include "b200_decode.h"
#include "host_utils.h"
#include "ptx.cuh"

namespace b200 {

static constexpr float FP8_MAX = 448.0f;

// ------------------------------------------------------------------------------------------------
// a-4  metadata.  Reference: 3rdparty/flashmla-fp8/csrc/kernels/get_mla_metadata.cu:8-75 (one warp, a serial walk over
// the sm parts on lane 0 with every operand fetched from global memory inside the loop).  Same integers here; the block
// counts and lengths are staged in shared memory by the whole CTA, and the walk -- sequential by construction: where a
// part ends decides where the next one starts -- runs on one lane out of registers, with the next request's block
// count loaded one step ahead so that no memory latency sits on the loop-carried chain.
// ------------------------------------------------------------------------------------------------
constexpr int kMetaThreads = 128;
__global__ void __launch_bounds__(kMetaThreads) mla_metadata_kernel(const int32_t* __restrict__ seqlens,
                                                                    int32_t* __restrict__ meta,
                                                                    int32_t* __restrict__ num_splits, int batch,
                                                                    int block_n, int fixed_overhead, int num_sm_parts) {
  extern __shared__ int32_t s_meta[];  // [batch] block counts, [batch] lengths, [warps] partial totals
  int32_t* s_blocks = s_meta;
  int32_t* s_len = s_meta + batch;
  int32_t* s_part = s_meta + 2 * batch;
  int total = 0;
  for (int i = threadIdx.x; i < batch; i += kMetaThreads) {
    const int len = seqlens[i];
    const int nb = (len + block_n - 1) / block_n;
    s_blocks[i] = nb;
    s_len[i] = len;
    total += nb + fixed_overhead;
  }
  for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x != 0) return;
  total = 0;
#pragma unroll
  for (int w = 0; w < kMetaThreads / 32; ++w) total += s_part[w];
  const int payload = max((total + num_sm_parts - 1) / num_sm_parts + fixed_overhead, 2 * fixed_overhead);
  int now_idx = 0, now_block = 0, now_split = 0, cum = 0;
  int nb = batch > 0 ? s_blocks[0] : 0, nb_next = batch > 1 ? s_blocks[1] : 0;  // requests now_idx, now_idx + 1
  int last_len = 0;                                                             // length of request now_idx - 1
  num_splits[0] = 0;
  for (int i = 0; i < num_sm_parts; ++i) {
    int4 head = make_int4(now_idx, now_block * block_n, 0, 0);
    const int begin_split = now_split;
    int remain = payload;
    while (now_idx < batch) {
      const int left = nb - now_block;
      if (remain >= left + fixed_overhead) {
        cum += now_split + 1;
        num_splits[now_idx + 1] = cum;
        remain -= left + fixed_overhead;
        last_len = s_len[now_idx];  // consumed at the end of the part at the earliest
        ++now_idx;
        nb = nb_next;
        nb_next = now_idx + 1 < batch ? s_blocks[now_idx + 1] : 0;
        now_block = 0;
        now_split = 0;
      } else {
        if (remain - fixed_overhead > 0) {
          now_block += remain - fixed_overhead;
          ++now_split;
          remain = 0;
        }
        break;
      }
    }
    head.z = now_block > 0 ? now_idx : now_idx - 1;
    head.w = now_block > 0 ? now_block * block_n : last_len;
    int4* m = reinterpret_cast<int4*>(meta + i * 8);
    m[0] = head;
    m[1] = make_int4(begin_split, 0, 0, 0);
  }
}

// ------------------------------------------------------------------------------------------------
// a-6 / a-7  per-row quantiser (one warp per row of 512 nope + 64 rope bf16 values).
// Reference arithmetic: cache_kernels.cu:27-98 / :140-218.
// ------------------------------------------------------------------------------------------------
// nope_in: 512 bf16 latent dims; rope_raw: this lane's two rope dims (2*lane, 2*lane+1) as packed bf16 bits.
__device__ __forceinline__ void quantize_row_parts(const __nv_bfloat16* __restrict__ nope_in, uint32_t rope_raw,
                                                   uint8_t* __restrict__ nope_out, float* __restrict__ scale_out,
                                                   __nv_bfloat16* __restrict__ rope_out, int lane) {
  // 512 bf16 = 1024 B = 2 x (32 lanes x 16 B)
  uint4 raw[2];
  raw[0] = __ldg(reinterpret_cast<const uint4*>(nope_in) + lane);
  raw[1] = __ldg(reinterpret_cast<const uint4*>(nope_in) + 32 + lane);
  float v[16];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + j * 2] = __uint_as_float(w[j] << 16);
      v[i * 8 + j * 2 + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) amax = fmaxf(amax, fabsf(v[i]));  // NaN-ignoring like __hmax
  amax = warp_max_f32(amax);
  float scale = __fdiv_rn(amax, FP8_MAX);
  scale = __fadd_rn(scale, 1e-12f);
  const float inv = __fdiv_rn(1.0f, scale);
  if (lane == 0) *scale_out = scale;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    uint2 packed;
    packed.x = cvt_e4m3x4(__fmul_rn(v[i * 8 + 0], inv), __fmul_rn(v[i * 8 + 1], inv), __fmul_rn(v[i * 8 + 2], inv),
                          __fmul_rn(v[i * 8 + 3], inv));
    packed.y = cvt_e4m3x4(__fmul_rn(v[i * 8 + 4], inv), __fmul_rn(v[i * 8 + 5], inv), __fmul_rn(v[i * 8 + 6], inv),
                          __fmul_rn(v[i * 8 + 7], inv));
    *reinterpret_cast<uint2*>(nope_out + (i * 32 + lane) * 8) = packed;
  }
  const float r0 = __fmul_rn(__uint_as_float(rope_raw << 16), inv);
  const float r1 = __fmul_rn(__uint_as_float(rope_raw & 0xFFFF0000u), inv);
  __nv_bfloat162 rr = __floats2bfloat162_rn(r0, r1);
  *reinterpret_cast<__nv_bfloat162*>(rope_out + lane * 2) = rr;
}

__device__ __forceinline__ void quantize_row_576(const __nv_bfloat16* __restrict__ in, uint8_t* __restrict__ nope_out,
                                                 float* __restrict__ scale_out, __nv_bfloat16* __restrict__ rope_out,
                                                 int lane) {
  quantize_row_parts(in, __ldg(reinterpret_cast<const uint32_t*>(in + 512) + lane), nope_out, scale_out, rope_out, lane);
}

// ------------------------------------------------------------------------------------------------
// the decode prologue in one launch.  For every token: rotary embedding of the 64 rope dims of each
// q head and of the new key (cos/sin from the model's cos_sin_cache row of the token's position), then the a-6 / a-7
// quantisers on the rotated rows -- q rows to (e4m3, scale, rope/scale), the key row scattered into the paged pool.
// Replaces apply_rope_with_cos_sin_cache_inplace (python/sglang/srt/layers/rotary_embedding.py:198-213, kernel
// 3rdparty/flashinfer/include/flashinfer/pos_enc.cuh:190-208) + MLATokenToKVPool.set_kv_buffer + the
// quantize_ckv_per_token_head call of the backend.  The rotation is x*cos + rot(x)*sin in fp32 with every operation
// rounded (the reference's DeepseekScalingRotaryEmbedding.forward_native, rotary_embedding.py:804-846, which the oracle
// is pinned against), the result is rounded to bf16 exactly where the unfused sequence stores bf16 -- so the outputs
// are bit-identical to rope -> bf16 -> quantiser.
// A lane owns rope dims (2*lane, 2*lane+1): one rotation pair when interleaved (GPT-J style, DeepSeek); for NeoX style
// (LongCat) the partner dims d +- 32 live in lane ^ 16.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t rope_pair(uint32_t raw, const float* __restrict__ cos_sin, bool neox, int lane) {
  const float x0 = __uint_as_float(raw << 16), x1 = __uint_as_float(raw & 0xFFFF0000u);
  float o0, o1;
  if (!neox) {
    const float c = __ldg(cos_sin + lane), s = __ldg(cos_sin + 32 + lane);
    o0 = __fadd_rn(__fmul_rn(x0, c), __fmul_rn(-x1, s));
    o1 = __fadd_rn(__fmul_rn(x1, c), __fmul_rn(x0, s));
  } else {
    const uint32_t other = __shfl_xor_sync(0xffffffffu, raw, 16);
    const float y0 = __uint_as_float(other << 16), y1 = __uint_as_float(other & 0xFFFF0000u);
    const int j = (2 * lane) & 31;  // cos/sin index of dim 2*lane (dims d and d+32 share one angle)
    const float2 c = __ldg(reinterpret_cast<const float2*>(cos_sin + j));
    const float2 s = __ldg(reinterpret_cast<const float2*>(cos_sin + 32 + j));
    const bool first_half = lane < 16;
    o0 = __fadd_rn(__fmul_rn(x0, c.x), __fmul_rn(first_half ? -y0 : y0, s.x));
    o1 = __fadd_rn(__fmul_rn(x1, c.y), __fmul_rn(first_half ? -y1 : y1, s.y));
  }
  const __nv_bfloat162 r = __floats2bfloat162_rn(o0, o1);
  return *reinterpret_cast<const uint32_t*>(&r);
}

struct RopeQuantParams {
  const __nv_bfloat16* q_nope;   // [tokens][heads][512], element strides below
  const __nv_bfloat16* q_pe;     // [tokens][heads][64], before rotation
  const __nv_bfloat16* kv;       // [tokens][576]: normed latent | k_pe before rotation
  int64_t q_nope_stride_t, q_nope_stride_h, q_pe_stride_t, q_pe_stride_h, kv_stride_t;
  const int64_t* positions;      // [tokens]
  const float* cos_sin_cache;    // [max_position][64]: cos[32] | sin[32]
  const int32_t* cache_loc;      // [tokens] pool slots; null: no key rows
  uint8_t* q_nope_out;           // [tokens][heads][512]
  float* q_scale_out;            // [tokens][heads]
  __nv_bfloat16* q_rope_out;     // [tokens][heads][64]
  uint8_t* k_lora_cache;
  float* k_scale_cache;
  __nv_bfloat16* k_rope_cache;
  int64_t tokens;
  int heads, neox;
};

__global__ void __launch_bounds__(256) rope_quantize_fused_kernel(const __grid_constant__ RopeQuantParams p) {
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t q_rows = p.tokens * p.heads;
  if (row < q_rows) {
    const int64_t t = row / p.heads;
    const int h = (int)(row - t * p.heads);
    const float* cs = p.cos_sin_cache + __ldg(p.positions + t) * 64;
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(p.q_pe + t * p.q_pe_stride_t + h * p.q_pe_stride_h) + lane);
    quantize_row_parts(p.q_nope + t * p.q_nope_stride_t + h * p.q_nope_stride_h, rope_pair(raw, cs, p.neox != 0, lane),
                       p.q_nope_out + row * 512, p.q_scale_out + row, p.q_rope_out + row * 64, lane);
  } else if (row < q_rows + (p.cache_loc ? p.tokens : 0)) {
    const int64_t t = row - q_rows;
    const float* cs = p.cos_sin_cache + __ldg(p.positions + t) * 64;
    const __nv_bfloat16* in = p.kv + t * p.kv_stride_t;
    const uint32_t raw = __ldg(reinterpret_cast<const uint32_t*>(in + 512) + lane);
    const int64_t loc = __ldg(p.cache_loc + t);
    quantize_row_parts(in, rope_pair(raw, cs, p.neox != 0, lane), p.k_lora_cache + loc * 512, p.k_scale_cache + loc,
                       p.k_rope_cache + loc * 64, lane);
  }
}

__global__ void __launch_bounds__(256) quantize_q_kernel(const __nv_bfloat16* __restrict__ q, uint8_t* __restrict__ nope,
                                                         float* __restrict__ scale, __nv_bfloat16* __restrict__ rope,
                                                         int64_t rows) {
  // launched with programmatic stream serialization: the kernel behind this one may start its prologue now, and this
  // one may have been scheduled while its predecessor drained -- griddepcontrol.wait orders the data
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  quantize_row_576(q + row * 576, nope + row * 512, scale + row, rope + row * 64, threadIdx.x & 31);
}

__global__ void __launch_bounds__(256)
    quantize_cache_k_kernel(const __nv_bfloat16* __restrict__ key, int64_t key_row_stride,
                            const int32_t* __restrict__ indices, uint8_t* __restrict__ lora_cache,
                            float* __restrict__ scale_cache, __nv_bfloat16* __restrict__ rope_cache, int64_t n) {
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int64_t loc = indices[row];
  quantize_row_576(key + row * key_row_stride, lora_cache + loc * 512, scale_cache + loc, rope_cache + loc * 64,
                   threadIdx.x & 31);
}

// ------------------------------------------------------------------------------------------------
// a-8  indexed de-quantisation.  Reference: cache_kernels.cu:274-369.  One warp per gathered token.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float e4m3_byte_to_float(uint32_t b) {
  __half_raw h = __nv_cvt_fp8_to_halfraw(static_cast<__nv_fp8_storage_t>(b), __NV_E4M3);
  return __half2float(__half(h));
}

__global__ void __launch_bounds__(256)
    dequantize_indexed_kernel(const uint8_t* __restrict__ lora_cache, const __nv_bfloat16* __restrict__ rope_cache,
                              const float* __restrict__ scale_cache, const int32_t* __restrict__ indices,
                              __nv_bfloat16* __restrict__ lora_out, __nv_bfloat16* __restrict__ rope_out, int64_t n) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const int64_t loc = indices[row];
    const float s = __ldg(scale_cache + loc);
    const uint4 raw = __ldg(reinterpret_cast<const uint4*>(lora_cache + loc * 512) + lane);  // 16 codes
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
    uint32_t o[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float f0 = __fmul_rn(e4m3_byte_to_float(w[j] & 0xFF), s);
      float f1 = __fmul_rn(e4m3_byte_to_float((w[j] >> 8) & 0xFF), s);
      float f2 = __fmul_rn(e4m3_byte_to_float((w[j] >> 16) & 0xFF), s);
      float f3 = __fmul_rn(e4m3_byte_to_float(w[j] >> 24), s);
      __nv_bfloat162 a = __floats2bfloat162_rn(f0, f1), b = __floats2bfloat162_rn(f2, f3);
      o[j * 2] = *reinterpret_cast<uint32_t*>(&a);
      o[j * 2 + 1] = *reinterpret_cast<uint32_t*>(&b);
    }
    uint4* dst = reinterpret_cast<uint4*>(lora_out + row * 512 + lane * 16);
    dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
    dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
    const uint32_t rr = __ldg(reinterpret_cast<const uint32_t*>(rope_cache + loc * 64) + lane);
    __nv_bfloat162 r = __floats2bfloat162_rn(__fmul_rn(__uint_as_float(rr << 16), s),
                                             __fmul_rn(__uint_as_float(rr & 0xFFFF0000u), s));
    *reinterpret_cast<__nv_bfloat162*>(rope_out + row * 64 + lane * 2) = r;
  }
}

// ------------------------------------------------------------------------------------------------
// a-5  split-KV combine.  Reference: ckv_fp8_per_token/mla_combine.cu:16-154.
// One warp per (request, q row): lse_g = log2(sum 2^lse_i), out = sum 2^(lse_i - lse_g) * oaccum_i.
// o_accum: [total_splits][M][512] fp32, lse_accum: [total_splits][M] fp32 (log2 domain).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    mla_combine_kernel(const float* __restrict__ o_accum, const float* __restrict__ lse_accum,
                       const int32_t* __restrict__ num_splits_ptr, __nv_bfloat16* __restrict__ out,
                       float* __restrict__ lse_out, int batch, int M, int h_q, int s_q, int merged_up_to) {
  pdl_wait_prior_grid();
  const int b = blockIdx.x;
  const int row = blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int s0 = num_splits_ptr[b], s1 = num_splits_ptr[b + 1];
  const int n = s1 - s0;
  // the decode kernel wrote the final result itself: an uncut request, or (pair kernel with a merge-state buffer) one
  // cut into at most `merged_up_to` pieces, which the part that finished last has merged
  if (n <= 1 || n <= merged_up_to) return;
  float max_lse = -INFINITY;
  for (int i = lane; i < n; i += 32) max_lse = fmaxf(max_lse, lse_accum[(size_t)(s0 + i) * M + row]);
  max_lse = warp_max_f32(max_lse);
  if (max_lse == -INFINITY) max_lse = 0.f;
  float sum = 0.f;
  for (int i = lane; i < n; i += 32) sum += exp2f(lse_accum[(size_t)(s0 + i) * M + row] - max_lse);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float global_lse = (sum == 0.f || sum != sum) ? INFINITY : log2f(sum) + max_lse;
  if (lane == 0) {
    // lse layout [b, h_q, s_q]; row = s * h_q + h
    const int s = row / h_q, h = row % h_q;
    lse_out[((size_t)b * h_q + h) * s_q + s] = global_lse * 0.6931471805599453f;
  }
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (int i = 0; i < n; ++i) {
    const float w = exp2f(lse_accum[(size_t)(s0 + i) * M + row] - global_lse);
    if (w == 0.f || w != w) continue;  // empty split: partial may hold anything
    const float4* src = reinterpret_cast<const float4*>(o_accum + ((size_t)(s0 + i) * M + row) * 512);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float4 v = __ldg(src + j * 32 + lane);
      acc[j * 4 + 0] += w * v.x;
      acc[j * 4 + 1] += w * v.y;
      acc[j * 4 + 2] += w * v.z;
      acc[j * 4 + 3] += w * v.w;
    }
  }
  __nv_bfloat16* dst = out + ((size_t)b * M + row) * 512;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 a = __floats2bfloat162_rn(acc[j * 4], acc[j * 4 + 1]);
    __nv_bfloat162 c = __floats2bfloat162_rn(acc[j * 4 + 2], acc[j * 4 + 3]);
    uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&c));
    *reinterpret_cast<uint2*>(dst + (j * 32 + lane) * 4) = pk;
  }
}

// ------------------------------------------------------------------------------------------------
// G-2  1x128 activation quantiser.  Reference: 3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-437 and
// python/sglang/test/test_block_fp8.py:15-40.  16 lanes per 128-element group (8 bf16 = 16 B per lane).
// scales are written column-major [k/128][scale_stride] (scale_stride = m padded to 4: TMA-aligned).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    act_quant_1x128_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_row_stride, int64_t x_group_stride,
                           uint8_t* __restrict__ q, float* __restrict__ scales, int64_t num_groups, int groups_per_row,
                           int64_t scale_stride, float eps) {
  // the GEMM behind this kernel is launched with programmatic stream serialization: let its CTAs start their prologue
  // (barrier init, TMEM allocation, descriptor prefetch) now; its griddepcontrol.wait still orders the data
  pdl_launch_dependents();
  pdl_wait_prior_grid();  // this kernel is itself launched as a programmatic dependent of its producer
  const int64_t g = (int64_t)blockIdx.x * (blockDim.x >> 4) + (threadIdx.x >> 4);
  const int sub = threadIdx.x & 15;
  if (g >= num_groups) return;
  // group (row, col) of the input sits at x + row * x_row_stride + col * x_group_stride (contiguous: k and 128): the
  // producer's output is read in place when it is a column slice ([m, k] of a wider row) or a per-head slice
  // ([m, heads, 128] of wider heads) -- no .contiguous() copy in front of the quantiser
  const int64_t row = g / groups_per_row, col = g % groups_per_row;
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(x + row * x_row_stride + col * x_group_stride) + sub);
  const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
  float v[8];
  float amax = eps;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[j * 2] = __uint_as_float(w[j] << 16);
    v[j * 2 + 1] = __uint_as_float(w[j] & 0xFFFF0000u);
    amax = fmaxf(amax, fmaxf(fabsf(v[j * 2]), fabsf(v[j * 2 + 1])));
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o, 16));
  const float s = __fdiv_rn(amax, FP8_MAX);
  if (sub == 0) scales[col * scale_stride + row] = s;
  float qv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) qv[j] = fminf(fmaxf(__fdiv_rn(v[j], s), -FP8_MAX), FP8_MAX);
  uint2 pk;
  pk.x = cvt_e4m3x4(qv[0], qv[1], qv[2], qv[3]);
  pk.y = cvt_e4m3x4(qv[4], qv[5], qv[6], qv[7]);
  *reinterpret_cast<uint2*>(q + g * 128 + sub * 8) = pk;
}

// ------------------------------------------------------------------------------------------------
// f-2  RMSNorm of the two halves of the fused qkv_a projection output, with the 1x128 activation quantiser of the
// q_b_proj input fused in.  Reference: FusedRMSNorm.forward (python/sglang/srt/layers/layernorm.py:276-303 ->
// flashinfer _rmsnorm_fused_parallel, 3rdparty/flashinfer/include/flashinfer/norm.cuh:474-630: out = bf16(float(x) *
// rsqrt(mean(x^2) + eps) * float(w))) called at models/deepseek_v2.py:816-820 (q_a -> a contiguous buffer, kv_a in
// place), and the block-quant epilogue of forward_with_allgather_fusion (fuse_block_quant_fp8, layernorm.py:305-330;
// block_scale plumbing deepseek_v2.py:796-807).  One CTA per token: warp w < q_dim / 512 owns columns [512 w, 512 w +
// 512) of q_a (four quantiser groups, 8 lanes x 16 values each), the last warp owns the 512 columns of kv_a.  The
// quantised codes are computed from the bf16-rounded normalised values with the arithmetic of act_quant_1x128_kernel,
// so they are bit-identical to quantising the bf16 output in a second launch.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
    qk_rmsnorm_quant_kernel(__nv_bfloat16* __restrict__ qkv, int64_t row_stride, int q_dim, const __nv_bfloat16* __restrict__ q_weight,
                            const __nv_bfloat16* __restrict__ kv_weight, float eps, __nv_bfloat16* __restrict__ q_out,
                            int64_t q_out_stride, uint8_t* __restrict__ q8, float* __restrict__ q_scales,
                            int64_t scale_stride, float quant_eps) {
  extern __shared__ float s_norm[];  // [warps] partial sums of squares
  pdl_launch_dependents();
  pdl_wait_prior_grid();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_warps = q_dim / 512;
  const bool is_kv = warp == q_warps;
  const int64_t row = blockIdx.x;
  __nv_bfloat16* src = qkv + row * row_stride + (is_kv ? q_dim : warp * 512) + lane * 16;
  const __nv_bfloat16* wsrc = (is_kv ? kv_weight : q_weight + warp * 512) + lane * 16;
  float v[16], w[16];
  {
    const uint4 r0 = *reinterpret_cast<const uint4*>(src), r1 = *reinterpret_cast<const uint4*>(src + 8);
    const uint4 w0 = __ldg(reinterpret_cast<const uint4*>(wsrc)), w1 = __ldg(reinterpret_cast<const uint4*>(wsrc + 8));
    const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    const uint32_t ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[2 * j] = __uint_as_float(rr[j] << 16), v[2 * j + 1] = __uint_as_float(rr[j] & 0xFFFF0000u);
      w[2 * j] = __uint_as_float(ww[j] << 16), w[2 * j + 1] = __uint_as_float(ww[j] & 0xFFFF0000u);
    }
  }
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) ss = __fmaf_rn(v[j], v[j], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) s_norm[warp] = ss;
  __syncthreads();
  float total = 0.f;
  if (is_kv) {
    total = s_norm[q_warps];
  } else {
    for (int i = 0; i < q_warps; ++i) total += s_norm[i];  // fixed order
  }
  const float rcp = rsqrtf(__fadd_rn(__fdiv_rn(total, is_kv ? 512.f : (float)q_dim), eps));
  uint32_t packed[8];
  float o[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(__fmul_rn(__fmul_rn(v[2 * j], rcp), w[2 * j]),
                                                    __fmul_rn(__fmul_rn(v[2 * j + 1], rcp), w[2 * j + 1]));
    packed[j] = *reinterpret_cast<const uint32_t*>(&h);
    o[2 * j] = __uint_as_float(packed[j] << 16), o[2 * j + 1] = __uint_as_float(packed[j] & 0xFFFF0000u);
  }
  if (is_kv) {  // kv_a: normalised in place (the k_nope columns the fused prologue quantises into the pool)
    *reinterpret_cast<uint4*>(src) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    *reinterpret_cast<uint4*>(src + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
    return;
  }
  const int col0 = warp * 512 + lane * 16;
  if (q_out != nullptr) {
    __nv_bfloat16* dst = q_out + row * q_out_stride + col0;
    *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
    *reinterpret_cast<uint4*>(dst + 8) = make_uint4(packed[4], packed[5], packed[6], packed[7]);
  }
  if (q8 != nullptr) {  // 1x128 groups = 8 lanes x 16 values; arithmetic of act_quant_1x128_kernel
    float amax = quant_eps;
#pragma unroll
    for (int j = 0; j < 16; ++j) amax = fmaxf(amax, fabsf(o[j]));
#pragma unroll
    for (int d = 4; d > 0; d >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, d, 8));
    const float sc = __fdiv_rn(amax, FP8_MAX);
    const int group = col0 >> 7;
    if ((lane & 7) == 0) q_scales[group * scale_stride + row] = sc;
    float qv[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) qv[j] = fminf(fmaxf(__fdiv_rn(o[j], sc), -FP8_MAX), FP8_MAX);
    uint4 pk;
    pk.x = cvt_e4m3x4(qv[0], qv[1], qv[2], qv[3]);
    pk.y = cvt_e4m3x4(qv[4], qv[5], qv[6], qv[7]);
    pk.z = cvt_e4m3x4(qv[8], qv[9], qv[10], qv[11]);
    pk.w = cvt_e4m3x4(qv[12], qv[13], qv[14], qv[15]);
    *reinterpret_cast<uint4*>(q8 + row * q_dim + col0) = pk;
  }
}

}  // namespace b200

using namespace b200;

// Split-KV combine launch; requests cut into at most `merged_up_to` pieces are skipped (merged by the decode kernel).
int launch_mla_combine(const float* o_accum, const float* lse_accum, const int32_t* num_splits, void* out, float* lse,
                       int batch, int s_q, int h_q, int merged_up_to, int use_pdl, void* stream) {
  if (batch == 0) return 0;  // IDLE batch
  B200_CHECK(o_accum && lse_accum && num_splits && out && lse, "combine: null pointer");
  const int M = s_q * h_q;
  const int warps = 8;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(batch, (M + warps - 1) / warps, 1);
  cfg.blockDim = dim3(warps * 32, 1, 1);
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  B200_CUDA(cudaLaunchKernelEx(&cfg, mla_combine_kernel, o_accum, lse_accum, num_splits, (__nv_bfloat16*)out, lse, batch,
                               M, h_q, s_q, merged_up_to));
  return 0;
}

extern "C" {

const char* b200_last_error(void) { return get_last_error(); }

int b200_sm_count(void) { return device_sm_count(); }

int b200_mla_num_sm_parts(int num_heads_per_head_k, int num_heads_k, int sm_count) {
  if (num_heads_per_head_k <= 0 || num_heads_k <= 0 || sm_count <= 0) return 0;
  const int m_blocks = (num_heads_per_head_k + B200_MLA_BLOCK_M - 1) / B200_MLA_BLOCK_M;
  return sm_count / num_heads_k / m_blocks;
}

int b200_mla_get_metadata(const int32_t* cache_seqlens, int batch, int num_sm_parts, int32_t* tile_scheduler_metadata,
                          int32_t* num_splits, void* stream) {
  return b200_mla_get_metadata_ex(cache_seqlens, batch, num_sm_parts, B200_MLA_FIXED_OVERHEAD_BLOCKS,
                                  tile_scheduler_metadata, num_splits, stream);
}

int b200_mla_get_metadata_ex(const int32_t* cache_seqlens, int batch, int num_sm_parts, int fixed_overhead_blocks,
                             int32_t* tile_scheduler_metadata, int32_t* num_splits, void* stream) {
  B200_CHECK((cache_seqlens || batch == 0) && tile_scheduler_metadata && num_splits, "get_metadata: null pointer");
  // batch == 0 is an IDLE batch (a data-parallel attention rank with no request, flash_api.cpp:27-63 accepts it): every
  // part then starts at request 0 >= batch, which the decode kernels read as "no work", and num_splits = [0]
  B200_CHECK(batch >= 0 && num_sm_parts > 0, "get_metadata: batch=%d num_sm_parts=%d", batch, num_sm_parts);
  B200_CHECK(fixed_overhead_blocks > 0 && fixed_overhead_blocks <= 1024, "get_metadata: fixed_overhead_blocks=%d",
             fixed_overhead_blocks);
  B200_CHECK((reinterpret_cast<uintptr_t>(tile_scheduler_metadata) & 15) == 0, "get_metadata: metadata must be 16-byte aligned");
  B200_CHECK((size_t)batch * 8 <= 200 * 1024, "get_metadata: batch %d too large", batch);
  const size_t smem = (size_t)batch * 8 + (kMetaThreads / 32) * 4;
  if (smem > 48 * 1024)
    B200_CUDA(cudaFuncSetAttribute(mla_metadata_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  mla_metadata_kernel<<<1, kMetaThreads, smem, (cudaStream_t)stream>>>(cache_seqlens, tile_scheduler_metadata, num_splits,
                                                                       batch, B200_MLA_PAGE_SIZE, fixed_overhead_blocks,
                                                                       num_sm_parts);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_mla_quantize_q(const void* q, int64_t rows, int head_dim, int head_dim_v, void* q_nope, float* q_scale,
                        void* q_rope, void* stream) {
  B200_CHECK(head_dim == 576 && head_dim_v == 512, "quantize_q: only head_dim=576/head_dim_v=512 supported (got %d/%d)",
             head_dim, head_dim_v);
  B200_CHECK(q && q_nope && q_scale && q_rope, "quantize_q: null pointer");
  if (rows == 0) return 0;
  const int warps = 8;
  return launch_kernel(quantize_q_kernel, dim3((unsigned)((rows + warps - 1) / warps)), dim3(warps * 32), 0,
                       (cudaStream_t)stream, pdl_default(), (const __nv_bfloat16*)q, (uint8_t*)q_nope, q_scale,
                       (__nv_bfloat16*)q_rope, rows);
}

int b200_mla_quantize_and_cache_k(const void* key, int64_t key_row_stride, int64_t n, int head_dim, int head_dim_v,
                                  const int32_t* indices, void* k_lora_cache, float* k_scale_cache, void* k_rope_cache,
                                  void* stream) {
  B200_CHECK(head_dim == 576 && head_dim_v == 512,
             "quantize_and_cache_k: only head_dim=576/head_dim_v=512 supported (got %d/%d)", head_dim, head_dim_v);
  B200_CHECK(key && indices && k_lora_cache && k_scale_cache && k_rope_cache, "quantize_and_cache_k: null pointer");
  B200_CHECK(key_row_stride >= 576 && key_row_stride % 8 == 0, "quantize_and_cache_k: key row stride %lld invalid",
             (long long)key_row_stride);
  if (n == 0) return 0;
  const int warps = 8;
  return launch_kernel(quantize_cache_k_kernel, dim3((unsigned)((n + warps - 1) / warps)), dim3(warps * 32), 0,
                       (cudaStream_t)stream, pdl_default(), (const __nv_bfloat16*)key, key_row_stride, indices,
                       (uint8_t*)k_lora_cache, k_scale_cache, (__nv_bfloat16*)k_rope_cache, n);
}

int b200_mla_rope_quantize_fused(const void* q_nope, int64_t q_nope_stride_t, int64_t q_nope_stride_h, const void* q_pe,
                                 int64_t q_pe_stride_t, int64_t q_pe_stride_h, const void* kv, int64_t kv_stride_t,
                                 const int64_t* positions, const float* cos_sin_cache, int rotary_dim, int is_neox,
                                 int64_t tokens, int heads, void* q_nope_out, float* q_scale_out, void* q_rope_out,
                                 const int32_t* cache_loc, void* k_lora_cache, float* k_scale_cache, void* k_rope_cache,
                                 void* stream) {
  B200_CHECK(rotary_dim == 64, "rope_quantize_fused: only rotary_dim=64 supported (got %d)", rotary_dim);
  B200_CHECK(q_nope && q_pe && positions && cos_sin_cache && q_nope_out && q_scale_out && q_rope_out,
             "rope_quantize_fused: null pointer");
  B200_CHECK(tokens >= 0 && heads > 0, "rope_quantize_fused: tokens=%lld heads=%d", (long long)tokens, heads);
  B200_CHECK(q_nope_stride_t % 8 == 0 && q_nope_stride_h % 8 == 0 && q_nope_stride_h >= 512 &&
                 ((uintptr_t)q_nope & 15) == 0,
             "rope_quantize_fused: q_nope rows must be 16-byte aligned (strides %lld, %lld)",
             (long long)q_nope_stride_t, (long long)q_nope_stride_h);
  B200_CHECK(q_pe_stride_t % 2 == 0 && q_pe_stride_h % 2 == 0 && q_pe_stride_h >= 64 && ((uintptr_t)q_pe & 3) == 0,
             "rope_quantize_fused: q_pe rows must be 4-byte aligned (strides %lld, %lld)", (long long)q_pe_stride_t,
             (long long)q_pe_stride_h);
  B200_CHECK(((uintptr_t)cos_sin_cache & 7) == 0, "rope_quantize_fused: cos_sin_cache must be 8-byte aligned");
  if (cache_loc) {
    B200_CHECK(kv && k_lora_cache && k_scale_cache && k_rope_cache, "rope_quantize_fused: null key / pool pointer");
    B200_CHECK(kv_stride_t >= 576 && kv_stride_t % 8 == 0 && ((uintptr_t)kv & 15) == 0,
               "rope_quantize_fused: kv row stride %lld invalid", (long long)kv_stride_t);
  }
  if (tokens == 0) return 0;
  RopeQuantParams prm;
  prm.q_nope = (const __nv_bfloat16*)q_nope;
  prm.q_pe = (const __nv_bfloat16*)q_pe;
  prm.kv = (const __nv_bfloat16*)kv;
  prm.q_nope_stride_t = q_nope_stride_t;
  prm.q_nope_stride_h = q_nope_stride_h;
  prm.q_pe_stride_t = q_pe_stride_t;
  prm.q_pe_stride_h = q_pe_stride_h;
  prm.kv_stride_t = kv_stride_t;
  prm.positions = positions;
  prm.cos_sin_cache = cos_sin_cache;
  prm.cache_loc = cache_loc;
  prm.q_nope_out = (uint8_t*)q_nope_out;
  prm.q_scale_out = q_scale_out;
  prm.q_rope_out = (__nv_bfloat16*)q_rope_out;
  prm.k_lora_cache = (uint8_t*)k_lora_cache;
  prm.k_scale_cache = k_scale_cache;
  prm.k_rope_cache = (__nv_bfloat16*)k_rope_cache;
  prm.tokens = tokens;
  prm.heads = heads;
  prm.neox = is_neox;
  const int warps = 8;
  const int64_t rows = tokens * heads + (cache_loc ? tokens : 0);
  return launch_kernel(rope_quantize_fused_kernel, dim3((unsigned)((rows + warps - 1) / warps)), dim3(warps * 32), 0,
                       (cudaStream_t)stream, pdl_default(), prm);
}

int b200_mla_dequantize_indexed(const void* k_lora_cache, const void* k_rope_cache, const float* k_scale_cache,
                                const int32_t* indices, int64_t n, int64_t max_num_tokens, void* k_lora_out,
                                void* k_rope_out, void* stream) {
  if (n == 0) return 0;  // empty request: the (empty) output tensors have no storage
  B200_CHECK(k_lora_cache && k_rope_cache && k_scale_cache && k_lora_out && k_rope_out, "dequantize: null pointer");
  B200_CHECK(indices, "dequantize: null indices");
  // grid depends only on max_num_tokens when given, so CUDA-graph capture and replay agree
  // (reference: cache_kernels.cu:372-423)
  const int warps = 8;
  int64_t sizing = max_num_tokens > 0 ? max_num_tokens : n;
  int64_t blocks = (sizing + warps - 1) / warps;
  if (blocks > 65535) blocks = 65535;
  dequantize_indexed_kernel<<<(unsigned)blocks, warps * 32, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)k_lora_cache, (const __nv_bfloat16*)k_rope_cache, k_scale_cache, indices,
      (__nv_bfloat16*)k_lora_out, (__nv_bfloat16*)k_rope_out, n);
  B200_CUDA(cudaGetLastError());
  return 0;
}

int b200_mla_combine(const float* o_accum, const float* lse_accum, const int32_t* num_splits, void* out, float* lse,
                     int batch, int s_q, int h_q, int use_pdl, void* stream) {
  return launch_mla_combine(o_accum, lse_accum, num_splits, out, lse, batch, s_q, h_q, 0, use_pdl, stream);
}

int b200_act_quant_1x128(const void* x, int64_t m, int64_t k, void* q, float* scales, int64_t scale_stride, float eps,
                         void* stream) {
  return b200_act_quant_1x128_strided(x, k, 128, m, k, q, scales, scale_stride, eps, stream);
}

int b200_act_quant_1x128_strided(const void* x, int64_t x_row_stride, int64_t x_group_stride, int64_t m, int64_t k,
                                 void* q, float* scales, int64_t scale_stride, float eps, void* stream) {
  B200_CHECK(x && q && scales, "act_quant: null pointer");
  B200_CHECK(k % 128 == 0, "act_quant: k=%lld must be a multiple of 128", (long long)k);
  B200_CHECK(scale_stride >= m, "act_quant: scale_stride %lld < m %lld", (long long)scale_stride, (long long)m);
  B200_CHECK(x_row_stride % 8 == 0 && x_group_stride % 8 == 0 && x_group_stride >= 128 && ((uintptr_t)x & 15) == 0,
             "act_quant: input groups must be 16-byte aligned (row stride %lld, group stride %lld)",
             (long long)x_row_stride, (long long)x_group_stride);
  if (m == 0) return 0;
  const int64_t groups = m * (k / 128);
  const int groups_per_block = 16;
  return launch_kernel(act_quant_1x128_kernel, dim3((unsigned)((groups + groups_per_block - 1) / groups_per_block)),
                       dim3(groups_per_block * 16), 0, (cudaStream_t)stream, pdl_default(), (const __nv_bfloat16*)x,
                       x_row_stride, x_group_stride, (uint8_t*)q, scales, groups, (int)(k / 128), scale_stride, eps);
}

int b200_qk_rmsnorm_quant(void* qkv, int64_t m, int64_t row_stride, int q_dim, int kv_dim, const void* q_weight,
                          const void* kv_weight, float eps, void* q_out, int64_t q_out_stride, void* q8, float* q_scales,
                          int64_t scale_stride, float quant_eps, void* stream) {
  B200_CHECK(qkv && q_weight && kv_weight, "qk_rmsnorm_quant: null pointer");
  B200_CHECK(kv_dim == 512 && q_dim > 0 && q_dim % 512 == 0 && q_dim <= 7 * 512,
             "qk_rmsnorm_quant: q_dim=%d (multiple of 512, <= 3584) kv_dim=%d (512) not supported", q_dim, kv_dim);
  B200_CHECK(q_out || q8, "qk_rmsnorm_quant: neither a bf16 nor a quantised q output was given");
  B200_CHECK(!q8 || q_scales, "qk_rmsnorm_quant: quantised output needs a scale buffer");
  B200_CHECK(row_stride % 8 == 0 && (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (!q_out || q_out_stride % 8 == 0),
             "qk_rmsnorm_quant: rows must be 16-byte aligned");
  if (m == 0) return 0;
  const int warps = q_dim / 512 + 1;
  return launch_kernel(qk_rmsnorm_quant_kernel, dim3((unsigned)m), dim3(warps * 32), warps * sizeof(float),
                       (cudaStream_t)stream, pdl_default(), (__nv_bfloat16*)qkv, row_stride, q_dim,
                       (const __nv_bfloat16*)q_weight, (const __nv_bfloat16*)kv_weight, eps, (__nv_bfloat16*)q_out,
                       q_out_stride, (uint8_t*)q8, q_scales, scale_stride, quant_eps);
}

}  // extern "C"
