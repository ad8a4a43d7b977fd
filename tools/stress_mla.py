"""Timing-perturbation stress: the MLA kernels must give bit-identical results while other streams hammer HBM / the SMs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import test_gpu_parity as T
from sglang_fluentllm_b200 import flash_mla_fp8 as fm
DEV = T.DEV
for (b, h, s_q, seqlens, causal) in ((2, 128, 1, [9000, 4100], False), (40, 128, 1, [700 + 37 * i for i in range(40)], False),
                                     (128, 128, 1, [4096] * 128, False), (64, 16, 1, [3000 + 7 * i for i in range(64)], False),
                                     (3, 128, 2, [3000, 130, 1500], True)):
    c = T.make_case(fm, b, h, s_q, seqlens, nan_pad=False)
    sl = c["sl"].to(DEV)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, fixed_overhead_blocks=fm.tuned_overhead_blocks(s_q * h))
    a = (c["qn"], c["qr"], c["kl"], c["kr"], c["qs"], c["ks"], c["bt"].to(DEV), sl, 512, meta, ns)
    o0, l0 = fm.flash_mla_ckv_fp8_per_token(*a, causal=causal, softmax_scale=192 ** -0.5)
    torch.cuda.synchronize()
    o0, l0 = o0.clone(), l0.clone()
    side = torch.cuda.Stream()
    big = torch.empty(1 << 28, dtype=torch.uint8, device=DEV)
    x = torch.randn(8192, 8192, device=DEV, dtype=torch.bfloat16)
    bad = 0
    for mode in ("quiet", "memcpy", "gemm", "both"):
        for it in range(40):
            with torch.cuda.stream(side):
                if mode in ("memcpy", "both"):
                    for _ in range(2):
                        big[: 1 << 27].copy_(big[1 << 27:])
                if mode in ("gemm", "both"):
                    y = x @ x
            o, l = fm.flash_mla_ckv_fp8_per_token(*a, causal=causal, softmax_scale=192 ** -0.5)
            torch.cuda.synchronize()
            if not (torch.equal(o, o0) and torch.equal(l, l0)):
                bad += 1
                print("MISMATCH", (b, h, s_q), mode, it, float((o.float() - o0.float()).abs().max()), float((l - l0).abs().max()))
    print("case b=%d h=%d s_q=%d: %d mismatching runs of 160" % (b, h, s_q, bad))
