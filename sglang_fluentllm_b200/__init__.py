"""B200-native (sm_100a) quantised decode hot path for SGLang-FluentLLM.

Sub-modules mirror the reference's names for this path:
  flash_mla_fp8      -- kernel-module API of 3rdparty/flashmla-fp8 (paged MLA decode over the per-token-FP8 KV pool)
  deep_gemm          -- deep_gemm.gemm_fp8_fp8_bf16_nt + layout helpers
  fp8_kernel         -- sglang_per_token_group_quant_fp8 (1x128 activation quantiser)
  fp8_linear         -- Fp8Config / Fp8LinearMethod / w8a8_block_fp8_linear operator surface
  flashmla_backend   -- FlashMLABackend (AttentionBackend plug-in) + the minimal pool / allocator it reads
  rope               -- fused decode prologue: rotary embedding + q quantiser + quantise-and-cache of the new key
  layernorm          -- FusedRMSNorm of q_a / kv_a, optionally emitting the q_b_proj input block-quantised
  install            -- sys.modules redirections of INTEGRATION.md (flash_mla_fp8 / flash_mla_swap / deep_gemm)
  tp                 -- one-box tensor-parallel harness (column/row-parallel FP8 linears + one NCCL all-reduce)
The CUDA code lives in csrc/ behind the C ABI declared in include/b200_decode.h; there is no CPU fallback.
"""
__version__ = "0.1.0"

from . import deep_gemm, flash_mla_fp8, flashmla_backend, fp8_kernel, fp8_linear, install, layernorm, rope, tp  # noqa: E402,F401
