// TEST INFRASTRUCTURE ONLY: C entry points over the reference's kernels compiled for the host (see ref_shim.h,
// build_ref.py).  Launch configurations are the reference's launchers':
//   get_mla_metadata.cu:82-86  <<<1, 32, (2 * batch + 1) ints>>>
//   cache_kernels.cu:111-121, 235-249  4 warps per block, one row per warp
//   cache_kernels.cu:388-423  4 warps per block, one query per block (queries_per_block when the grid is capped)
//   sgl_fp8_quantization.cu:452-491  16 threads per 128-element group, up to 16 groups per block
#include "shim/ref_shim_impl.inc"

int shared_mem[1 << 20];  // get_mla_metadata_kernel's dynamic shared memory (extern __shared__ int shared_mem[])

extern "C" {

int ref_get_mla_metadata(const int* seqlens, int batch, int block_size_n, int fixed_overhead_num_blocks, int num_sm_parts,
                         int* tile_scheduler_metadata /* [num_sm_parts][8], 16-byte aligned */, int* num_splits) {
  if (2 * batch + 1 > (1 << 20)) return 1;
  Mla_metadata_params p;
  p.seqlens_k_ptr = const_cast<int*>(seqlens);
  p.tile_scheduler_metadata_ptr = tile_scheduler_metadata;
  p.num_splits_ptr = num_splits;
  p.batch_size = batch;
  p.block_size_n = block_size_n;
  p.fixed_overhead_num_blocks = fixed_overhead_num_blocks;
  p.num_sm_parts = num_sm_parts;
  refshim::run_grid(1, 32, [&] { get_mla_metadata_kernel(p); });
  return 0;
}

int ref_quantize_ckv_per_token_head(const uint16_t* in_bf16, int rows, int head_dim, int head_dim_v, uint8_t* nope,
                                    float* scale, uint16_t* rope_bf16) {
  const int warp_num = 4;
  refshim::run_grid((rows + warp_num - 1) / warp_num, warp_num * 32, [&] {
    quantize_single_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(in_bf16),
                                          reinterpret_cast<__nv_fp8_e4m3*>(nope), scale,
                                          reinterpret_cast<__nv_bfloat16*>(rope_bf16), rows, head_dim, head_dim_v);
  });
  return 0;
}

int ref_quantize_and_cache_k(const uint16_t* key_bf16, const int* loc, uint8_t* k_lora, float* k_scale, uint16_t* k_rope_bf16,
                             int n, int lora_stride, int scale_stride, int rope_stride, int head_dim, int head_dim_v) {
  const int warp_num = 4;
  refshim::run_grid((n + warp_num - 1) / warp_num, warp_num * 32, [&] {
    quant_and_save_cache_k_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_bfloat16*>(key_bf16), loc,
                                                 reinterpret_cast<__nv_fp8_e4m3*>(k_lora), k_scale,
                                                 reinterpret_cast<__nv_bfloat16*>(k_rope_bf16), n, lora_stride, scale_stride,
                                                 rope_stride, head_dim, head_dim_v);
  });
  return 0;
}

int ref_dequantize_fused_indexed(const uint8_t* k_lora, const uint16_t* k_rope_bf16, const float* k_scale, const int* indices,
                                 uint16_t* lora_out_bf16, uint16_t* rope_out_bf16, int num_queries, int max_num_tokens,
                                 int lora_stride, int scale_stride, int rope_stride) {
  // cache_kernels.cu:388-408
  const int MAX_BLOCK_NUM = 65535;
  int block_num, queries_per_block;
  const int required = num_queries <= MAX_BLOCK_NUM ? num_queries : MAX_BLOCK_NUM;
  if (max_num_tokens > 0 && max_num_tokens < required) {
    block_num = max_num_tokens;
    queries_per_block = (num_queries + max_num_tokens - 1) / max_num_tokens;
  } else {
    block_num = required;
    queries_per_block = num_queries <= MAX_BLOCK_NUM ? 1 : (num_queries + MAX_BLOCK_NUM - 1) / MAX_BLOCK_NUM;
  }
  refshim::run_grid(block_num, 4 * 32, [&] {
    dequantize_fused_indexed_kernel<__nv_bfloat16>(reinterpret_cast<const __nv_fp8_e4m3*>(k_lora),
                                                   reinterpret_cast<const __nv_bfloat16*>(k_rope_bf16), k_scale, indices,
                                                   reinterpret_cast<__nv_bfloat16*>(lora_out_bf16),
                                                   reinterpret_cast<__nv_bfloat16*>(rope_out_bf16), num_queries,
                                                   lora_stride, scale_stride, rope_stride, 512, 64, queries_per_block);
  });
  return 0;
}

int ref_per_token_group_quant_fp8(const uint16_t* x_bf16, int64_t m, int64_t k, uint8_t* q, float* scales,
                                  int64_t scale_stride, float eps) {
  // sgl_fp8_quantization.cu:452-491 (group 128, column-major float scales)
  const int group_size = 128;
  const int64_t num_groups = m * k / group_size;
  int groups_per_block = 1;
  if (num_groups % 16 == 0) groups_per_block = 16;
  else if (num_groups % 8 == 0) groups_per_block = 8;
  else if (num_groups % 4 == 0) groups_per_block = 4;
  else if (num_groups % 2 == 0) groups_per_block = 2;
  refshim::run_grid((unsigned)(num_groups / groups_per_block), (unsigned)(groups_per_block * 16), [&] {
    per_token_group_quant_8bit_kernel<__nv_bfloat16, __nv_fp8_e4m3, true, false>(
        reinterpret_cast<const __nv_bfloat16*>(x_bf16), q, scales, group_size, (int)num_groups, groups_per_block, eps,
        -448.0f, 448.0f, (int)(k / group_size), (int)scale_stride);
  });
  return 0;
}

}  // extern "C"
