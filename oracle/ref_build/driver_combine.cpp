// TEST INFRASTRUCTURE ONLY: C entry point over the reference's split-KV combine kernel compiled for the host
// (ckv_fp8_per_token/mla_combine.cu:16-154; launch shape of its launcher, :180-203: grid (b, ceil(M / 8)), 8 warps,
// MAX_SPLITS = num_sm_parts rounded up to 32).
namespace ckv_fp8_per_token {
float smem_buf[8 * (160 + 1)];  // the kernel's `extern __shared__ float smem_buf[]`: BLOCK_SIZE_M x (MAX_SPLITS + 1)
}

using ckv_fp8_per_token::Flash_fwd_mla_params;

template <int MAX_SPLITS>
static void ref_combine_run(Flash_fwd_mla_params& p, int batch, int M) {
  constexpr int BLOCK_SIZE_M = 8, NUM_THREADS = BLOCK_SIZE_M * 32;
  refshim::run_grid((unsigned)batch, NUM_THREADS, [&] {
    ckv_fp8_per_token::flash_fwd_mla_combine_kernel<cutlass::bfloat16_t, 512, BLOCK_SIZE_M, MAX_SPLITS, NUM_THREADS>(p);
  }, (unsigned)((M + BLOCK_SIZE_M - 1) / BLOCK_SIZE_M));
}

extern "C" int ref_mla_combine(const float* o_accum, const float* lse_accum, const int* num_splits, uint16_t* out_bf16,
                               float* lse, int batch, int s_q, int h_q, int num_sm_parts) {
  // out bf16 [b][s_q][h_q][512] contiguous, lse f32 [b][M] (flash_api.cpp:285-292, 330-345 with h_k = 1)
  Flash_fwd_mla_params p;
  std::memset(&p, 0, sizeof(p));
  const int M = s_q * h_q;
  p.b = batch;
  p.s_q = s_q;
  p.q_seq_per_hk = M;
  p.h_q = h_q;
  p.h_k = 1;
  p.q_head_per_hk = h_q;
  p.d = 576;
  p.d_v = 512;
  p.num_sm_parts = num_sm_parts;
  p.num_splits_ptr = const_cast<int*>(num_splits);
  p.softmax_lseaccum_ptr = const_cast<float*>(lse_accum);
  p.oaccum_ptr = const_cast<float*>(o_accum);
  p.softmax_lse_ptr = lse;
  p.o_ptr = out_bf16;
  p.o_batch_stride = (int64_t)M * 512;
  p.o_head_stride = 512;
  p.o_row_stride = 512;
  if (num_sm_parts <= 32) ref_combine_run<32>(p, batch, M);
  else if (num_sm_parts <= 64) ref_combine_run<64>(p, batch, M);
  else if (num_sm_parts <= 96) ref_combine_run<96>(p, batch, M);
  else if (num_sm_parts <= 128) ref_combine_run<128>(p, batch, M);
  else if (num_sm_parts <= 160) ref_combine_run<160>(p, batch, M);
  else return 1;
  return 0;
}
