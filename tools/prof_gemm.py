"""Minimal launcher for ncu: a few block-FP8 GEMM launches at one shape.  Usage: python tools/prof_gemm.py M N K"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sglang_fluentllm_b200 import deep_gemm as dg, fp8_kernel as fk
from sglang_fluentllm_b200.fp8_linear import dummy_block_fp8_weight
m, n, k = (int(x) for x in sys.argv[1:4])
dev = torch.device("cuda:0"); torch.manual_seed(0)
ws = [dummy_block_fp8_weight(n, k, dev) for _ in range(4)]
a8, a_s = fk.sglang_per_token_group_quant_fp8(torch.randn(m, k, dtype=torch.bfloat16, device=dev), 128, column_major_scales=True, scale_tma_aligned=True)
out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
for i in range(6):
    dg.gemm_fp8_fp8_bf16_nt((a8, a_s), ws[i % 4], out, True)
torch.cuda.synchronize()
