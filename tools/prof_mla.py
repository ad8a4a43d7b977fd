"""Minimal launcher for ncu: a few MLA decode launches at one shape.  Usage: python tools/prof_mla.py H [B SEQ]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sglang_fluentllm_b200 import flash_mla_fp8 as fm
h = int(sys.argv[1]); b = int(sys.argv[2]) if len(sys.argv) > 2 else 128; seq = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda:0"); torch.manual_seed(0)
pp = seq // 64 + 1; n_pages = b * pp + 1
kv = torch.randn(n_pages * 64, 1, 576, dtype=torch.bfloat16, device=dev)
kl = torch.empty(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev); ks = torch.empty(n_pages * 64, 1, 1, device=dev); kr = torch.empty(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=dev), 512)
q = torch.randn(b, 1, h, 576, dtype=torch.bfloat16, device=dev)
qn, qs, qr = fm.quantize_ckv_per_token_head(q, 512)
bt = (torch.randperm(n_pages - 1)[: b * pp].reshape(b, pp) + 1).to(torch.int32).to(dev)
sl = torch.full((b,), seq, dtype=torch.int32, device=dev)
meta, ns = fm.get_mla_metadata(sl, h, 1, fixed_overhead_blocks=fm.tuned_overhead_blocks(h))  # the backend's schedule
for _ in range(4):
    fm.flash_mla_ckv_fp8_per_token(qn, qr, kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), qs, ks.view(n_pages, 64, 1, 1), bt, sl, 512, meta, ns,
                                   softmax_scale=192 ** -0.5)  # the model's scale (logit spread 2.5 bits)
torch.cuda.synchronize()
