/*
 * b200_decode.h -- C ABI of the B200-native (sm_100a) quantised decode hot path:
 *   paged MLA decode attention over a per-token-FP8 KV cache + block-scaled W8A8 FP8 GEMM.
 *
 * Drop-in boundary for meituan-longcat/SGLang-FluentLLM.  Each entry point names the reference interface it
 * replaces (paths relative to the reference root).  Conventions:
 *   - all pointers are DEVICE pointers unless stated; the caller owns every buffer (inputs, outputs, scratch), so
 *     nothing is allocated here and every call is CUDA-graph capturable;
 *   - `stream` is a cudaStream_t passed as void*;
 *   - return value 0 = ok; non-zero = error, message via b200_last_error() (thread-local).  Nothing exits or aborts
 *     the process (the reference's CHECK_CUDA does, csrc/kernels/ckv_fp8_per_token/utils.h:7-16);
 *   - no global mutable state besides per-device attribute caches.
 */
#ifndef B200_DECODE_H_
#define B200_DECODE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_MLA_PAGE_SIZE 64              /* flashmla-fp8 csrc/kernels/ckv_fp8_per_token/config.h:6 */
#define B200_MLA_BLOCK_M 64                /* config.h:5: q rows per CTA */
#define B200_MLA_FIXED_OVERHEAD_BLOCKS 5   /* config.h:14 */
#define B200_MLA_METADATA_INTS 8           /* params.h:72 TileSchedulerMetaDataSize */
#define B200_MLA_HEAD_DIM_V 512
#define B200_MLA_HEAD_DIM_ROPE 64

/* Last error message of the calling thread ("" if none). */
const char* b200_last_error(void);

/* Number of SMs of the current device (148 on B200). */
int b200_sm_count(void);

/* num_sm_parts = sm_count / num_heads_k / ceil(num_heads_per_head_k / 64).
 * Replaces the shape rule in 3rdparty/flashmla-fp8/csrc/flash_api.cpp:41-43. */
int b200_mla_num_sm_parts(int num_heads_per_head_k, int num_heads_k, int sm_count);

/* Tile-scheduler metadata.  Replaces flash_mla_fp8.get_mla_metadata
 * (flash_mla_fp8/flash_mla_interface.py:8-23 -> csrc/flash_api.cpp:27-63 -> csrc/kernels/get_mla_metadata.cu:8-82).
 *   cache_seqlens            int32 [batch]
 *   tile_scheduler_metadata  int32 [num_sm_parts][8]  {begin_idx, begin_seqlen, end_idx, end_seqlen, begin_n_split, 0,0,0}
 *   num_splits               int32 [batch + 1]        prefix sums of the per-request split counts */
int b200_mla_get_metadata(const int32_t* cache_seqlens, int batch, int num_sm_parts, int32_t* tile_scheduler_metadata,
                          int32_t* num_splits, void* stream);
/* The same scheduler with the per-request fixed cost as a parameter (get_mla_metadata.cu:16 fixes it at 5 blocks of 64
 * tokens, the cost of a request boundary on the reference's sm_90 kernel; here a boundary costs 12-24 blocks' worth of
 * time).  Same output format; any positive value gives a valid schedule for the decode kernels. */
int b200_mla_get_metadata_ex(const int32_t* cache_seqlens, int batch, int num_sm_parts, int fixed_overhead_blocks,
                             int32_t* tile_scheduler_metadata, int32_t* num_splits, void* stream);

/* Per-(token, head) FP8 quantiser for q.  Replaces flash_mla_fp8.quantize_ckv_per_token_head
 * (csrc/flash_api.cpp:399-460, cache_kernels.cu:27-122).
 *   q       bf16 [rows][576] contiguous      q_nope  e4m3 [rows][512]
 *   q_scale f32  [rows]                      q_rope  bf16 [rows][64]  (divided by the same scale) */
int b200_mla_quantize_q(const void* q, int64_t rows, int head_dim, int head_dim_v, void* q_nope, float* q_scale,
                        void* q_rope, void* stream);

/* Quantise new K tokens and scatter them into the paged pool.  Replaces flash_mla_fp8.quantize_and_cache_k
 * (csrc/flash_api.cpp:462-532, cache_kernels.cu:140-250).
 *   key bf16 [n][key_row_stride >= 576]; indices int32 [n] token slots;
 *   pool: k_lora_cache e4m3 [slots][512], k_scale_cache f32 [slots], k_rope_cache bf16 [slots][64]
 *   (the MLATokenToKVPool per_token_head layout, python/sglang/srt/mem_cache/memory_pool.py:635-649). */
int b200_mla_quantize_and_cache_k(const void* key, int64_t key_row_stride, int64_t n, int head_dim, int head_dim_v,
                                  const int32_t* indices, void* k_lora_cache, float* k_scale_cache, void* k_rope_cache,
                                  void* stream);

/* Decode prologue in one launch: rotary embedding of q_pe and of the new key's rope dims, the
 * per-(token, head) FP8 quantiser of the absorbed q, and quantise + scatter of the new key into the paged pool.
 * Replaces the sequence the reference runs per layer: apply_rope_with_cos_sin_cache_inplace
 * (python/sglang/srt/layers/rotary_embedding.py:198-213 with output_q_rope / fused_set_kv_buffer_arg as wired in
 * python/sglang/srt/models/deepseek_v2.py:843-857; kernel 3rdparty/flashinfer/include/flashinfer/pos_enc.cuh:190-208),
 * MLATokenToKVPool.set_kv_buffer -> quantize_and_cache_k (mem_cache/memory_pool.py:854-871) and
 * quantize_ckv_per_token_head (layers/attention/flashmla_backend.py:205-206).  Bit-identical to that sequence with
 * the rotation evaluated as in DeepseekScalingRotaryEmbedding.forward_native (rotary_embedding.py:804-846).
 *   q_nope bf16 [tokens][heads][512] and q_pe bf16 [tokens][heads][64] (before rotation): element strides given,
 *     last dimension contiguous, q_nope rows 16-byte aligned;
 *   kv bf16 [tokens][576] = normed latent | k_pe before rotation, row stride kv_stride_t;
 *   positions int64 [tokens]; cos_sin_cache f32 [max_position][64] = cos[32] | sin[32]; rotary_dim must be 64;
 *   is_neox 0: rotation pairs (2i, 2i+1) (DeepSeek), 1: pairs (i, i+32) (LongCat);
 *   outputs as b200_mla_quantize_q (contiguous [tokens][heads][...]) and, when cache_loc != NULL, the pool tensors of
 *   b200_mla_quantize_and_cache_k written at slots cache_loc int32 [tokens]. */
int b200_mla_rope_quantize_fused(const void* q_nope, int64_t q_nope_stride_t, int64_t q_nope_stride_h, const void* q_pe,
                                 int64_t q_pe_stride_t, int64_t q_pe_stride_h, const void* kv, int64_t kv_stride_t,
                                 const int64_t* positions, const float* cos_sin_cache, int rotary_dim, int is_neox,
                                 int64_t tokens, int heads, void* q_nope_out, float* q_scale_out, void* q_rope_out,
                                 const int32_t* cache_loc, void* k_lora_cache, float* k_scale_cache, void* k_rope_cache,
                                 void* stream);

/* Gather + de-quantise pool entries.  Replaces flash_mla_fp8.dequantize_ckv_fused_indexed
 * (csrc/flash_api.cpp:534-632, cache_kernels.cu:274-423).  max_num_tokens > 0 fixes the grid for CUDA graphs.
 *   k_lora_out bf16 [n][512], k_rope_out bf16 [n][64] */
int b200_mla_dequantize_indexed(const void* k_lora_cache, const void* k_rope_cache, const float* k_scale_cache,
                                const int32_t* indices, int64_t n, int64_t max_num_tokens, void* k_lora_out,
                                void* k_rope_out, void* stream);

/* Scratch the decode call needs: o_accum f32 [(batch + num_sm_parts)][M][512] then lse_accum f32 [(..)][M]
 * (same sizing as csrc/flash_api.cpp:365-368); M = s_q * h_q. */
size_t b200_mla_workspace_bytes(int batch, int s_q, int h_q, int num_sm_parts);

/* Paged MLA decode attention over the per-token-FP8 KV pool (split-KV + combine).
 * Replaces flash_mla_fp8.flash_mla_ckv_fp8_per_token (flash_mla_interface.py:68-120 ->
 * csrc/flash_api.cpp:216-397 -> ckv_fp8_per_token/splitkv_mla.cu:1147-1515 + mla_combine.cu:16-154).
 *   q_nope e4m3 [b][s_q][h_q][512]   q_rope bf16 [b][s_q][h_q][64]   q_scale f32 [b][s_q][h_q]   (contiguous)
 *   k_lora e4m3 [num_pages*64][512]  k_rope bf16 [num_pages*64][64]  k_scale f32 [num_pages*64]  (token-major pool)
 *   block_table int32 [b][block_table_stride]   cache_seqlens int32 [b]
 *   tile_scheduler_metadata / num_splits from b200_mla_get_metadata (num_sm_parts rows)
 *   out bf16 [b][s_q][h_q][512]      lse f32 [b][h_q][s_q]
 *   workspace >= b200_mla_workspace_bytes(...)
 *   merge_state: NULL, or int32 [b200_mla_merge_state_ints(batch, s_q, h_q)] that is ZERO when the call starts and is
 *   not shared with a call running concurrently on another stream.  With it, the kernel for s_q * h_q > 32 merges the
 *   requests the scheduler cut over several parts itself (the part that finishes last adds the others' partials) and
 *   no combine kernel is launched; the kernel leaves the buffer zero again, so one buffer zeroed once serves every
 *   later call on the stream.  NULL: partials + b200_mla_combine, as the reference. */
size_t b200_mla_merge_state_ints(int batch, int s_q, int h_q);
int b200_mla_decode_ckv_fp8_per_token(const void* q_nope, const void* q_rope, const float* q_scale, const void* k_lora,
                                      const void* k_rope, const float* k_scale, int64_t num_pages,
                                      const int32_t* block_table, int64_t block_table_stride,
                                      const int32_t* cache_seqlens, const int32_t* tile_scheduler_metadata,
                                      const int32_t* num_splits, int num_sm_parts, int batch, int s_q, int h_q,
                                      float softmax_scale, int is_causal, void* out, float* lse, void* workspace,
                                      size_t workspace_bytes, int32_t* merge_state, int use_pdl, void* stream);

/* Paged MLA decode attention over a bf16 KV cache (split-KV + combine), no quantisation anywhere.
 * Replaces flash_mla_fp8.flash_mla_with_kvcache / flash_mla_swap.flash_mla_with_kvcache
 * (flash_mla_fp8/flash_mla_interface.py:26-66 -> csrc/flash_api.cpp:65-214 -> csrc/kernels/splitkv_mla.cu), the call
 * the backend makes when the pool is not per-token FP8 (python/sglang/srt/layers/attention/flashmla_backend.py:240-256).
 *   q bf16 [b][s_q][h_q][576] contiguous   k_cache bf16 [num_pages*64][576] (token-major; V = first 512 dims)
 *   other arguments, outputs and workspace as b200_mla_decode_ckv_fp8_per_token. */
int b200_mla_decode_bf16(const void* q, const void* k_cache, int64_t num_pages, const int32_t* block_table,
                         int64_t block_table_stride, const int32_t* cache_seqlens,
                         const int32_t* tile_scheduler_metadata, const int32_t* num_splits, int num_sm_parts,
                         int batch, int s_q, int h_q, float softmax_scale, int is_causal, void* out, float* lse,
                         void* workspace, size_t workspace_bytes, int use_pdl, void* stream);

/* Split-KV combine (also run by the decode call).  Replaces ckv_fp8_per_token/mla_combine.cu:16-154. */
int b200_mla_combine(const float* o_accum, const float* lse_accum, const int32_t* num_splits, void* out, float* lse,
                     int batch, int s_q, int h_q, int use_pdl, void* stream);

/* ---- one-box collectives over NVLink peer memory (csrc/p2p_comm.cu) -------------------------------------------------
 * Replace, for the exchanges either side of the attention block: tensor_model_parallel_all_reduce behind the
 * row-parallel o_proj (python/sglang/srt/layers/linear.py:1139) and the uneven token reduce-scatter / all-gather of the
 * attention-TP group (distributed/decoder_comm_manager.py:42-89; sm_90 implementation:
 * distributed/device_communicators/custom_triton_rsag/triton_rsag.py:97-300).
 * Every rank allocates one communication buffer (buf_bytes, the same on every rank) and one flag pad
 * (b200_comm_flag_bytes) with b200_comm_alloc, exports both (64-byte CUDA IPC handles), and opens its peers' (same
 * node).  bufs / flags: HOST arrays [world] of the device pointers as mapped in this process (own entries = own
 * allocations).  A call is collective: every rank of the group launches it with the same sizes (`blocks` is reserved:
 * pass 0) on its own operand `in`; the kernel stages the operand in HALF of the buffer (the halves alternate per call, which is
 * what lets a call end without a barrier), so an operand may take at most buf_bytes / 2.  bf16, 16-byte granularity;
 * sums in rank order in fp32 (every rank gets the same bits).  The pads carry the call counters: only these entry
 * points may touch them, and a call that trapped (a peer never arrived) leaves the group unusable. */
size_t b200_comm_flag_bytes(void);
int b200_comm_alloc(size_t bytes, void** ptr);
int b200_comm_free(void* ptr);
int b200_comm_ipc_export(void* ptr, void* handle64);
int b200_comm_ipc_open(const void* handle64, void** ptr);
int b200_comm_ipc_close(void* ptr);
/* out[numel] = sum over ranks of in[numel] (two ranks: one pass over the peer's operand; more: two-shot). */
int b200_comm_all_reduce_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world, const void* in,
                              int64_t numel, void* out, int blocks, void* stream);
/* The same sum through the switch (NVLS).  multicast_ptr / local_ptr: the multicast and the local mapping of ONE
 * symmetric buffer of buf_bytes per rank bound to a multicast object spanning the group (e.g. what
 * torch.distributed._symmetric_memory's rendezvous returns); flags: the pads of this library as above. */
int b200_comm_all_reduce_nvls_bf16(void* multicast_ptr, void* local_ptr, void* const* flags, size_t buf_bytes, int rank,
                                   int world, const void* in, int64_t numel, void* out, int blocks, void* stream);
/* row_offsets: HOST int64 [world + 1], first row of every rank's share.  in [row_offsets[world]][hidden] on every
 * rank, out [rows of this rank][hidden]. */
int b200_comm_reduce_scatter_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world,
                                  const int64_t* row_offsets, const void* in, int64_t hidden, void* out, int blocks,
                                  void* stream);
/* in [rows of this rank][hidden], out [row_offsets[world]][hidden] = every rank's rows in rank order. */
int b200_comm_all_gather_bf16(void* const* bufs, void* const* flags, size_t buf_bytes, int rank, int world,
                              const int64_t* row_offsets, const void* in, int64_t hidden, void* out, int blocks,
                              void* stream);

/* 1x128 activation quantiser, column-major TMA-aligned scales.  Replaces sgl_per_token_group_quant_fp8
 * (3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-520; python wrapper
 * python/sglang/srt/layers/dense/gemms/fp8/fp8_kernel.py:430-463).
 *   x bf16 [m][k]; q e4m3 [m][k]; scales f32 [k/128][scale_stride], scale_stride = m rounded up to 4. */
int b200_act_quant_1x128(const void* x, int64_t m, int64_t k, void* q, float* scales, int64_t scale_stride, float eps,
                         void* stream);
/* Same quantiser reading its input in place from a strided producer output: group (i, g) of the logical [m][k] input
 * is the 128 contiguous bf16 values at x + i * x_row_stride + g * x_group_stride (elements; both multiples of 8).
 * Covers a column slice of a wider row (row stride > k, group stride 128: the q_a columns of the fused qkv_a projection,
 * python/sglang/srt/models/deepseek_v2.py:796-807) and a per-head slice [m][heads][128] of wider heads (group stride =
 * head stride).  q is written contiguous [m][k]. */
int b200_act_quant_1x128_strided(const void* x, int64_t x_row_stride, int64_t x_group_stride, int64_t m, int64_t k,
                                 void* q, float* scales, int64_t scale_stride, float eps, void* stream);

/* RMSNorm of the fused qkv_a projection output (q_a columns [0, q_dim), kv_a columns [q_dim, q_dim + 512) of each row)
 * with the 1x128 activation quantiser of the q_b_proj input fused in.  Replaces FusedRMSNorm.forward
 * (python/sglang/srt/layers/layernorm.py:276-303 -> flashinfer _rmsnorm_fused_parallel,
 * 3rdparty/flashinfer/include/flashinfer/norm.cuh:474-630) and the fuse_block_quant_fp8 epilogue of its allgather-fused
 * form (layernorm.py:305-330), called at python/sglang/srt/models/deepseek_v2.py:796-820.
 *   qkv bf16 [m][row_stride]: kv_a is normalised in place; q_a is written normalised to q_out bf16 [m][q_out_stride]
 *   (may be NULL) and / or quantised to q8 e4m3 [m][q_dim] + q_scales f32 [q_dim/128][scale_stride] (may be NULL), the
 *   same bits b200_act_quant_1x128 produces from q_out.  q_dim: multiple of 512, <= 3584; kv_dim must be 512. */
int b200_qk_rmsnorm_quant(void* qkv, int64_t m, int64_t row_stride, int q_dim, int kv_dim, const void* q_weight,
                          const void* kv_weight, float eps, void* q_out, int64_t q_out_stride, void* q8, float* q_scales,
                          int64_t scale_stride, float quant_eps, void* stream);

/* Block-scaled W8A8 FP8 GEMM, out[m][n] = bf16( sum_kb a_scales[kb][i] * b_scales[n/128][kb] * (A_kb . B_kb^T) ).
 * Replaces deep_gemm.gemm_fp8_fp8_bf16_nt (3rdparty/deep_gemm/deep_gemm/jit_kernels/gemm.py:145-226 ->
 * include/deep_gemm/fp8_gemm.cuh:60-399,409-827); same argument order as the JIT'd
 * `launch(lhs, lhs_scales, rhs, rhs_scales, out, m, stream, ...)` C symbol (deep_gemm/jit/template.py:57-79).
 *   a e4m3 [m][k]; a_scales f32 [k/128][a_scale_stride] (column-major, stride = m rounded up to 4);
 *   b e4m3 [n][k]; b_scales f32 [ceil(n/128)][k/128]; out bf16 [m][n].  Requires n % 64 == 0, k % 128 == 0.
 *   workspace: optional split-K scratch of b200_fp8_gemm_workspace_bytes(m, n, k) bytes (fp32 partial slabs, one per
 *   split; no initialisation needed, one buffer serves every call on a stream).  NULL = no split-K. */
size_t b200_fp8_gemm_workspace_bytes(int m, int n, int k);
int b200_fp8_gemm_nt(const void* a, const float* a_scales, int64_t a_scale_stride, const void* b, const float* b_scales,
                     void* out, int m, int n, int k, void* workspace, size_t workspace_bytes, int use_pdl,
                     void* stream);

/* Grouped block-scaled FP8 GEMMs (MoE expert GEMMs).  Replaces deep_gemm.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous /
 * _masked / _offset (3rdparty/deep_gemm/deep_gemm/jit_kernels/m_grouped_gemm.py:55-127 / :130-215 / :225-310).
 *   b e4m3 [num_groups][n][k]; b_scales f32 [num_groups][ceil(n/128)][k/128]; n % 128 == 0.
 *   mode 1 (contiguous): a e4m3 [m][k], out bf16 [m][n], m % 128 == 0, group_layout = m_indices int32 [m]: every
 *     128-row block carries one group id, -1 skips the block; a_scales f32 [k/128][a_scale_stride] as b200_fp8_gemm_nt.
 *   mode 2 (masked): a e4m3 [num_groups][m][k], out bf16 [num_groups][m][n] (m = rows per group, a multiple of 128 when
 *     num_groups > 1), group_layout = masked_m int32 [num_groups]: rows of each group that are computed; a_scales f32
 *     [num_groups][k/128][a_scale_stride], transposed per group.
 *   mode 3 (offset): a e4m3 [m][k], out bf16 [m][n], any m, group_layout = m_offset int32 [num_groups + 1]
 *     (non-decreasing, m_offset[num_groups] <= m): group g owns rows [m_offset[g], m_offset[g + 1]); rows no group
 *     owns are left untouched; a_scales as mode 1. */
int b200_fp8_gemm_grouped_row_tiles(int mode, int num_groups, int m); /* grid.y of the launch (host-only helper) */
int b200_fp8_gemm_grouped_nt(const void* a, const float* a_scales, int64_t a_scale_stride, const void* b,
                             const float* b_scales, void* out, const int32_t* group_layout, int mode, int num_groups,
                             int m, int n, int k, int use_pdl, void* stream);

/* Diagnostic: clock64 pipeline trace of the last CTA-pair MLA launch made with B200_MLA_TRACE set. */
int b200_mla_read_trace(unsigned long long* out, int n);
/* Diagnostic: per (part, CTA rank) globaltimer ns at entry / at the end of the softmax warps' work, same launch. */
int b200_mla_read_part_times(unsigned long long* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* B200_DECODE_H_ */
