"""Block-scaled FP8 GEMM: this repo's kernel against the Blackwell kernel the reference ships
(3rdparty/deep_gemm_oss, sm100_fp8_gemm_1d2d via fp8_gemm_nt(..., disable_ue8m0_cast=True): fp32 scales, the same
1x128 / 128x128 recipe), installed into baseline/_ref by
    pip install --no-index --no-build-isolation --no-deps --target baseline/_ref <copy of 3rdparty/deep_gemm_oss>
on identical inputs, CUDA-graph replay over rotated weights (L2-cold), plus the FP8 / bf16 dense peaks of the box
(torch._scaled_mm / torch.matmul 8192^3, burst and a 3 s sustained loop) that tensor-pipe fractions are quoted against.
Usage: python tools/gemm_compare.py [--no-ref]   (GPU box; deep_gemm_oss JIT-compiles its kernels with nvcc on first use)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from sglang_fluentllm_b200 import deep_gemm as dg
from sglang_fluentllm_b200 import fp8_kernel as fk
from sglang_fluentllm_b200.fp8_linear import dummy_block_fp8_weight

dev = torch.device("cuda:0")
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = PEAKS.get("hbm_gbs", 6650.0)


def time_graph(fn, n_rot, reps=5):
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(min(n_rot, 3)):
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


def dense_peak(dtype_name):
    n = 8192
    if dtype_name == "fp8":
        a = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn)
        b = torch.randn(n, n, device=dev).to(torch.float8_e4m3fn).t()
        one = torch.ones((), device=dev)
        f = lambda: torch._scaled_mm(a, b, scale_a=one, scale_b=one, out_dtype=torch.bfloat16)
    else:
        a = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
        b = torch.randn(n, n, device=dev, dtype=torch.bfloat16)
        f = lambda: torch.matmul(a, b)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2 * n ** 3 / (best * 1e-3) / 1e12
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    k = 0
    while time.time() - t0 < 3.0:
        for _ in range(20):
            f()
        k += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    sustained = 2 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    return burst, sustained


def main():
    use_ref = "--no-ref" not in sys.argv
    ref = None
    if use_ref:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        os.environ.setdefault("DG_JIT_CACHE_DIR", "/tmp/dg_oss_cache")
        try:
            import deep_gemm_oss as ref
        except Exception as e:  # say exactly why
            print("deep_gemm_oss unavailable: %r" % (e,), flush=True)
            ref = None
    out = {"peaks": {}, "gemm": []}
    for name in ("fp8", "bf16"):
        try:
            burst, sus = dense_peak(name)
            out["peaks"][name] = {"burst_tflops": round(burst, 1), "sustained_tflops": round(sus, 1)}
            print("dense %s 8192^3: burst %.0f TFLOP/s, sustained (3 s) %.0f TFLOP/s" % (name, burst, sus), flush=True)
        except Exception as e:
            print("dense %s peak failed: %r" % (name, e), flush=True)
    fp8_peak = out["peaks"].get("fp8", {}).get("burst_tflops")
    for (n, k) in ((7168, 2048), (4096, 7168), (2112, 7168), (24576, 1536), (7168, 16384)):
        for m in (64, 128, 256, 4096):
            torch.manual_seed(0)
            n_rot = max(2, min(24, int(3e8 // (n * k)) + 2))
            ws = [dummy_block_fp8_weight(n, k, dev) for _ in range(n_rot)]
            x = torch.randn(m, k, dtype=torch.bfloat16, device=dev)
            a8, a_s = fk.sglang_per_token_group_quant_fp8(x, 128, column_major_scales=True, scale_tma_aligned=True)
            d = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
            ours = time_graph(lambda i: dg.gemm_fp8_fp8_bf16_nt((a8, a_s), ws[i % n_rot], d, True), n_rot)
            d_ours = d.clone()
            theirs = None
            if ref is not None:
                try:
                    d2 = torch.empty_like(d)
                    f = lambda i: ref.fp8_gemm_nt((a8, a_s), ws[i % n_rot], d2, disable_ue8m0_cast=True)
                    theirs = time_graph(f, n_rot)
                    dg.gemm_fp8_fp8_bf16_nt((a8, a_s), ws[(n_rot - 1) % n_rot], d, True)
                    ref.fp8_gemm_nt((a8, a_s), ws[(n_rot - 1) % n_rot], d2, disable_ue8m0_cast=True)
                    torch.cuda.synchronize()
                    diff = float((d.float() - d2.float()).norm() / d2.float().norm())
                except Exception as e:
                    print("deep_gemm_oss failed at m=%d n=%d k=%d: %r" % (m, n, k, e), flush=True)
                    theirs, diff = None, None
                    ref = None if "JIT" in repr(e) or "nvcc" in repr(e) else ref
            bytes_ = m * k + n * k + 2 * m * n + 4 * m * k // 128 + 4 * ((n + 127) // 128) * (k // 128)
            tf = 2.0 * m * n * k / ours / 1e6
            rec = {"m": m, "n": n, "k": k, "ours_us": round(ours, 2), "hbm_frac": round(bytes_ / ours / 1e3 / HBM, 3),
                   "tflops": round(tf, 1), "fp8_peak_frac": round(tf / fp8_peak, 3) if fp8_peak else None,
                   "deep_gemm_oss_us": round(theirs, 2) if theirs else None,
                   "rel_diff_vs_deep_gemm_oss": round(diff, 5) if theirs else None}
            out["gemm"].append(rec)
            print(json.dumps(rec), flush=True)
    print("SUMMARY " + json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
