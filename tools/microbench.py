"""Kernel microbenchmarks on one GPU (CUDA-graph replay, L2-cold by rotating over distinct buffers).
Usage: python tools/microbench.py [mla] [gemm]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sglang_fluentllm_b200 import flash_mla_fp8 as fm, deep_gemm as dg, fp8_kernel as fk, _lib
from sglang_fluentllm_b200.fp8_linear import dummy_block_fp8_weight

# the tile schedule the backend runs (per-request cost tuned for B200); a tree without the extension keeps the reference's
# softmax scale of the model (DeepSeek-V3: qk_head_dim ** -0.5 = 192 ** -0.5; on unit-variance q and k that is a logit
# spread of 2.5 in log2 units, which decides how often a row reference moves); B200_MB_SCALE overrides
SM_SCALE = float(os.environ.get("B200_MB_SCALE", 192 ** -0.5))
SCHED = lambda m: ({"fixed_overhead_blocks": fm.tuned_overhead_blocks(m), **({"num_sm_parts": fm.tuned_num_sm_parts(m)} if hasattr(fm, "tuned_num_sm_parts") and not os.environ.get("B200_REF_PARTS") else {})} if hasattr(fm, "tuned_overhead_blocks") else {})

dev = torch.device("cuda:0")
PEAK = 6555.8


def time_graph(fn, n_rot, reps=3):
    torch.cuda.synchronize()  # inputs were produced on the default stream
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(min(n_rot, 2)):
            fn(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_rot):
                fn(i)
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(reps):
            g.replay()
        e1.record(s)
        s.synchronize()
    return e0.elapsed_time(e1) / (reps * n_rot) * 1e3  # us


def bench_mla(b, seq, h, s_q=1, n_rot=None):
    torch.manual_seed(0)
    pages_per = (seq + 63) // 64 + 1
    n_pages = b * pages_per + 1
    kv_bytes = n_pages * 64 * 644
    if n_rot is None:
        n_rot = max(2, min(16, int(2.5e8 // kv_bytes) + 2))
    pools = []
    tmp = torch.empty(n_pages * 64, 1, 576, dtype=torch.bfloat16, device=dev)
    idx = torch.arange(n_pages * 64, dtype=torch.int32, device=dev)
    for _ in range(n_rot):
        tmp.normal_()
        kl = torch.empty(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev)
        ks = torch.empty(n_pages * 64, 1, 1, dtype=torch.float32, device=dev)
        kr = torch.empty(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
        fm.quantize_and_cache_k(tmp, kl, ks, kr, idx, 512)
        pools.append((kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), ks.view(n_pages, 64, 1, 1)))
    del tmp
    q = torch.randn(b, s_q, h, 576, dtype=torch.bfloat16, device=dev)
    qn, qs, qr = fm.quantize_ckv_per_token_head(q, 512)
    bt = (torch.randperm(n_pages - 1)[: b * pages_per].reshape(b, pages_per) + 1).to(torch.int32).to(dev)
    sl = torch.full((b,), seq, dtype=torch.int32, device=dev)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, **SCHED(s_q * h))
    L = _lib.lib()
    out = torch.empty(b, s_q, h, 512, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(b, h, s_q, dtype=torch.float32, device=dev)
    ws = torch.empty(L.b200_mla_workspace_bytes(b, s_q, h, meta.shape[0]), dtype=torch.uint8, device=dev)

    def fn(i):
        kl, kr, ks = pools[i % n_rot]
        fm.flash_mla_ckv_fp8_per_token(qn, qr, kl, kr, qs, ks, bt, sl, 512, meta, ns, softmax_scale=SM_SCALE,
                                       causal=s_q > 1, out=out, lse=lse, workspace=ws)

    us = time_graph(fn, n_rot)
    S, M = b * seq, h * s_q
    bytes_ = S * 644 + b * M * 644 + b * M * 1024
    flops = 2 * S * M * 1088
    print("mla b=%d seq=%d h=%d s_q=%d: %.1f us  %.0f GB/s (%.1f%% of %.0f)  %.0f TFLOP/s" % (
        b, seq, h, s_q, us, bytes_ / us / 1e3, 100 * bytes_ / us / 1e3 / PEAK, PEAK, flops / us / 1e6), flush=True)


def bench_mla_bf16(b, seq, h, s_q=1):
    """bf16-KV kernel (a-2): 1152 B per KV token."""
    torch.manual_seed(0)
    pages_per = (seq + 63) // 64 + 1
    n_pages = b * pages_per + 1
    n_rot = max(2, min(8, int(2.5e8 // (n_pages * 64 * 1152)) + 2))
    pools = [torch.randn(n_pages, 64, 1, 576, dtype=torch.bfloat16, device=dev) for _ in range(n_rot)]
    q = torch.randn(b, s_q, h, 576, dtype=torch.bfloat16, device=dev)
    bt = (torch.randperm(n_pages - 1)[: b * pages_per].reshape(b, pages_per) + 1).to(torch.int32).to(dev)
    sl = torch.full((b,), seq, dtype=torch.int32, device=dev)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, **SCHED(s_q * h))
    L = _lib.lib()
    out = torch.empty(b, s_q, h, 512, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(b, h, s_q, dtype=torch.float32, device=dev)
    ws = torch.empty(L.b200_mla_workspace_bytes(b, s_q, h, meta.shape[0]), dtype=torch.uint8, device=dev)

    def fn(i):
        fm.flash_mla_with_kvcache(q, pools[i % n_rot], bt, sl, 512, meta, ns, causal=s_q > 1, out=out, lse=lse,
                                  workspace=ws)

    us = time_graph(fn, n_rot)
    S, M = b * seq, h * s_q
    bytes_ = S * 1152 + b * M * 1152 + b * M * 1024
    print("mla-bf16 b=%d seq=%d h=%d s_q=%d: %.1f us  %.0f GB/s (%.1f%% of %.0f)  %.0f TFLOP/s" % (
        b, seq, h, s_q, us, bytes_ / us / 1e3, 100 * bytes_ / us / 1e3 / PEAK, PEAK, 2 * S * M * 1088 / us / 1e6),
        flush=True)


def bench_gemm(m, n, k, n_rot=None):
    torch.manual_seed(0)
    wbytes = n * k
    if n_rot is None:
        n_rot = max(2, min(24, int(3e8 // wbytes) + 2))
    ws_ = [dummy_block_fp8_weight(n, k, dev) for _ in range(n_rot)]
    x = torch.randn(m, k, dtype=torch.bfloat16, device=dev)
    a8, a_s = fk.sglang_per_token_group_quant_fp8(x, 128, column_major_scales=True, scale_tma_aligned=True)
    out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)

    def fn(i):
        dg.gemm_fp8_fp8_bf16_nt((a8, a_s), ws_[i % n_rot], out, True)

    us = time_graph(fn, n_rot)
    bytes_ = m * k + n * k + 2 * m * n + 4 * m * k // 128 + 4 * ((n + 127) // 128) * (k // 128)
    print("gemm m=%d n=%d k=%d: %.2f us  %.0f GB/s (%.1f%%)  %.0f TFLOP/s" % (
        m, n, k, us, bytes_ / us / 1e3, 100 * bytes_ / us / 1e3 / PEAK, 2.0 * m * n * k / us / 1e6), flush=True)


what = sys.argv[1:] or ["mla", "gemm"]
if what and what[0] == "gemm1":  # one shape: python tools/microbench.py gemm1 M N K  (with B200_GEMM_BN / B200_GEMM_SPLITS)
    bench_gemm(int(what[1]), int(what[2]), int(what[3]))
    what = []
if what and what[0] == "mla1":  # shapes: python tools/microbench.py mla1 B,SEQ,H[,SQ] ...
    for a in what[1:]:
        v = [int(x) for x in a.split(",")]
        bench_mla(v[0], v[1], v[2], s_q=v[3] if len(v) > 3 else 1)
    what = []
if "mla128" in what:
    bench_mla(128, 4096, 128)
    bench_mla(64, 4096, 128, s_q=2)
    bench_mla(128, 4096, 64, s_q=2)
if "mla" in what:
    bench_mla(128, 4096, 128)
    bench_mla(128, 4096, 16)
    bench_mla(128, 4096, 64)
    bench_mla(128, 4096, 32)
    bench_mla(64, 4096, 16, s_q=4)
    bench_mla(256, 8192, 16, n_rot=2)
if "bf16" in what:
    bench_mla_bf16(128, 4096, 16)
    bench_mla_bf16(128, 4096, 128)
    bench_mla_bf16(64, 4096, 16, s_q=4)
if "gemm" in what:
    for m in (16, 64, 128):
        bench_gemm(m, 7168, 2048)
        bench_gemm(m, 4096, 7168)
    bench_gemm(128, 2112, 7168)
    bench_gemm(128, 24576, 1536)
    bench_gemm(128, 7168, 16384)
    bench_gemm(4096, 7168, 2048)
    bench_gemm(4096, 4096, 7168)
