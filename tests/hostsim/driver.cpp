// TEST INFRASTRUCTURE ONLY: C entry points that run the kernels of csrc/small_kernels.cu on CPU threads with the grid /
// block shapes of the product's own launchers (same file, C-ABI section).
#include "../../oracle/ref_build/shim/ref_shim_impl.inc"

namespace b200 {
int32_t s_meta[1 << 20];  // mla_metadata_kernel's dynamic shared memory
float s_norm[64];         // qk_rmsnorm_quant_kernel's
}
using namespace b200;

extern "C" {

int host_mla_get_metadata(const int32_t* seqlens, int batch, int num_sm_parts, int32_t* meta, int32_t* num_splits) {
  refshim::run_grid(1, kMetaThreads, [&] {
    mla_metadata_kernel(seqlens, meta, num_splits, batch, B200_MLA_PAGE_SIZE, B200_MLA_FIXED_OVERHEAD_BLOCKS, num_sm_parts);
  });
  return 0;
}

int host_quantize_q(const uint16_t* q, int64_t rows, uint8_t* nope, float* scale, uint16_t* rope) {
  refshim::run_grid((unsigned)((rows + 7) / 8), 256, [&] {
    quantize_q_kernel(reinterpret_cast<const __nv_bfloat16*>(q), nope, scale, reinterpret_cast<__nv_bfloat16*>(rope), rows);
  });
  return 0;
}

int host_quantize_and_cache_k(const uint16_t* key, int64_t key_row_stride, int64_t n, const int32_t* indices, uint8_t* lora,
                              float* scale, uint16_t* rope) {
  refshim::run_grid((unsigned)((n + 7) / 8), 256, [&] {
    quantize_cache_k_kernel(reinterpret_cast<const __nv_bfloat16*>(key), key_row_stride, indices, lora, scale,
                            reinterpret_cast<__nv_bfloat16*>(rope), n);
  });
  return 0;
}

int host_dequantize_indexed(const uint8_t* lora, const uint16_t* rope, const float* scale, const int32_t* indices, int64_t n,
                            int64_t max_num_tokens, uint16_t* lora_out, uint16_t* rope_out) {
  int64_t sizing = max_num_tokens > 0 ? max_num_tokens : n;
  int64_t blocks = (sizing + 7) / 8;
  if (blocks > 65535) blocks = 65535;
  refshim::run_grid((unsigned)blocks, 256, [&] {
    dequantize_indexed_kernel(lora, reinterpret_cast<const __nv_bfloat16*>(rope), scale, indices,
                              reinterpret_cast<__nv_bfloat16*>(lora_out), reinterpret_cast<__nv_bfloat16*>(rope_out), n);
  });
  return 0;
}

int host_mla_combine(const float* o_accum, const float* lse_accum, const int32_t* num_splits, uint16_t* out, float* lse,
                     int batch, int s_q, int h_q) {
  const int M = s_q * h_q;
  refshim::run_grid((unsigned)batch, 256, [&] {
    mla_combine_kernel(o_accum, lse_accum, num_splits, reinterpret_cast<__nv_bfloat16*>(out), lse, batch, M, h_q, s_q, 0);
  }, (unsigned)((M + 7) / 8));
  return 0;
}

int host_act_quant_1x128_strided(const uint16_t* x, int64_t row_stride, int64_t group_stride, int64_t m, int64_t k,
                                 uint8_t* q, float* scales, int64_t scale_stride, float eps) {
  const int64_t groups = m * (k / 128);
  refshim::run_grid((unsigned)((groups + 15) / 16), 256, [&] {
    act_quant_1x128_kernel(reinterpret_cast<const __nv_bfloat16*>(x), row_stride, group_stride, q, scales, groups,
                           (int)(k / 128), scale_stride, eps);
  });
  return 0;
}

int host_rope_quantize_fused(const uint16_t* q_nope, int64_t qn_st, int64_t qn_sh, const uint16_t* q_pe, int64_t qp_st,
                             int64_t qp_sh, const uint16_t* kv, int64_t kv_st, const int64_t* positions,
                             const float* cos_sin_cache, int is_neox, int64_t tokens, int heads, uint8_t* q_nope_out,
                             float* q_scale_out, uint16_t* q_rope_out, const int32_t* cache_loc, uint8_t* k_lora,
                             float* k_scale, uint16_t* k_rope) {
  RopeQuantParams p;
  p.q_nope = reinterpret_cast<const __nv_bfloat16*>(q_nope);
  p.q_pe = reinterpret_cast<const __nv_bfloat16*>(q_pe);
  p.kv = reinterpret_cast<const __nv_bfloat16*>(kv);
  p.q_nope_stride_t = qn_st;
  p.q_nope_stride_h = qn_sh;
  p.q_pe_stride_t = qp_st;
  p.q_pe_stride_h = qp_sh;
  p.kv_stride_t = kv_st;
  p.positions = positions;
  p.cos_sin_cache = cos_sin_cache;
  p.cache_loc = cache_loc;
  p.q_nope_out = q_nope_out;
  p.q_scale_out = q_scale_out;
  p.q_rope_out = reinterpret_cast<__nv_bfloat16*>(q_rope_out);
  p.k_lora_cache = k_lora;
  p.k_scale_cache = k_scale;
  p.k_rope_cache = reinterpret_cast<__nv_bfloat16*>(k_rope);
  p.tokens = tokens;
  p.heads = heads;
  p.neox = is_neox;
  const int64_t rows = tokens * heads + (cache_loc ? tokens : 0);
  refshim::run_grid((unsigned)((rows + 7) / 8), 256, [&] { rope_quantize_fused_kernel(p); });
  return 0;
}

}  // extern "C"
