"""Grouped block-FP8 GEMMs on one GPU at MoE expert shapes: time, achieved TFLOP/s and weight-streaming GB/s.
CUDA-graph replay, CUDA events, L2-cold by rotating over distinct weight sets.  Usage: python tools/grouped_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sglang_fluentllm_b200 import deep_gemm as dg  # noqa: E402

dev = torch.device("cuda:0")


def time_graph(fn, n_rot, reps=5):
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


def fp8_rand(*shape):
    return (torch.randn(*shape, device=dev, dtype=torch.bfloat16)).to(torch.float8_e4m3fn)


def main():
    torch.manual_seed(0)
    for groups, m_per, n, k in ((32, 128, 4096, 7168), (32, 128, 7168, 2048), (8, 512, 4096, 7168), (8, 512, 7168, 2048)):
        w_bytes = groups * n * k
        n_rot = max(2, min(6, int(3e8 // w_bytes) + 2))
        ws = [(fp8_rand(groups, n, k), torch.rand(groups, n // 128, k // 128, device=dev) + 0.5) for _ in range(n_rot)]
        m = groups * m_per
        a = fp8_rand(m, k)
        a_s = dg.get_col_major_tma_aligned_tensor(torch.rand(m, k // 128, device=dev) + 0.5)
        out = torch.empty(m, n, dtype=torch.bfloat16, device=dev)
        idx = torch.arange(groups, dtype=torch.int32, device=dev).repeat_interleave(m_per)
        off = (torch.arange(groups + 1, dtype=torch.int32, device=dev) * m_per)
        a3, a3_s = a.view(groups, m_per, k), torch.rand(groups, m_per, k // 128, device=dev) + 0.5
        out3 = out.view(groups, m_per, n)
        masked = torch.full((groups,), m_per, dtype=torch.int32, device=dev)
        a3_s = dg.get_col_major_tma_aligned_tensor(a3_s)
        t_c = time_graph(lambda i: dg.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous((a, a_s), ws[i], out, idx, True), n_rot)
        t_m = time_graph(lambda i: dg.m_grouped_gemm_fp8_fp8_bf16_nt_masked((a3, a3_s), ws[i], out3, masked, m_per, True), n_rot)
        t_o = time_graph(lambda i: dg.m_grouped_gemm_fp8_fp8_bf16_nt_offset((a, a_s), ws[i], out, off, True), n_rot)
        flops, nbytes = 2.0 * m * n * k, w_bytes + m * k + 2 * m * n
        for name, t in (("contiguous", t_c), ("masked", t_m), ("offset", t_o)):
            print("grouped %-10s groups=%d m/group=%d n=%d k=%d: %.1f us  %.0f TFLOP/s  %.0f GB/s algorithmic" % (
                name, groups, m_per, n, k, t, flops / t / 1e6, nbytes / t / 1e3), flush=True)


if __name__ == "__main__":
    main()
