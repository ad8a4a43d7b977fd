"""Decode prologue on one GPU: the fused launch (b200_mla_rope_quantize_fused) against the sequence it replaces --
rope (torch elementwise, standing in for flashinfer's apply_rope_with_cos_sin_cache_inplace), the concat/.contiguous()
copy of Q, quantize_ckv_per_token_head and quantize_and_cache_k.  CUDA-graph replay, CUDA events, L2-cold by rotating
over distinct inputs.  Usage: python tools/prologue_bench.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from sglang_fluentllm_b200 import flash_mla_fp8 as fm, rope  # noqa: E402

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


def torch_rope(x, positions, cache):
    cs = cache[positions]
    cos, sin = cs[:, None, :32].repeat_interleave(2, -1), cs[:, None, 32:].repeat_interleave(2, -1)
    xf = x.float()
    rot = torch.stack([-xf[..., 1::2], xf[..., 0::2]], -1).flatten(-2)
    return (xf * cos + rot * sin).to(torch.bfloat16)


def main():
    torch.manual_seed(0)
    max_pos, slots, n_rot = 8192, 64 * 4096, 8
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float32) / 64))
    freqs = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    cache = torch.cat([freqs.cos(), freqs.sin()], -1).contiguous().to(dev)
    k_lora = torch.zeros(slots, 512, dtype=torch.uint8, device=dev)
    k_scale = torch.zeros(slots, dtype=torch.float32, device=dev)
    k_rope = torch.zeros(slots, 64, dtype=torch.bfloat16, device=dev)
    for tokens, heads in ((128, 128), (128, 16), (256, 16), (64 * 4, 16)):
        ins = []
        for _ in range(n_rot):
            q_buf = torch.randn(tokens, heads, 576, device=dev).to(torch.bfloat16)
            q_full = torch.randn(tokens, heads, 192, device=dev).to(torch.bfloat16)
            qkv_a = torch.randn(tokens, 2112, device=dev).to(torch.bfloat16)
            pos = torch.randint(0, max_pos, (tokens,), device=dev)
            loc = (torch.randperm(slots - 64)[:tokens] + 64).to(torch.int32).to(dev)
            ins.append((q_buf, q_full[..., 128:], qkv_a[:, 1536:], pos, loc))

        def fused(i):
            q_buf, q_pe, kv, pos, loc = ins[i]
            arg = rope.FusedSetKVBufferArg(value=kv[:, :512], k_buffer=k_rope, v_buffer=k_lora, scale_buffer=k_scale,
                                           cache_loc=loc)
            rope.fused_rope_quantize_set_kv(pos, q_buf[..., :512], q_pe, kv[:, 512:], cache, False, arg)

        def unfused(i):
            q_buf, q_pe, kv, pos, loc = ins[i]
            q_buf[..., 512:] = torch_rope(q_pe, pos, cache)
            key = torch.cat([kv[:, :512], torch_rope(kv[:, None, 512:], pos, cache)[:, 0]], -1)
            fm.quantize_and_cache_k(key, k_lora, k_scale, k_rope, loc, 512)
            fm.quantize_ckv_per_token_head(q_buf.contiguous(), 512)

        def quantisers_only(i):
            q_buf, q_pe, kv, pos, loc = ins[i]
            fm.quantize_and_cache_k(kv, k_lora, k_scale, k_rope, loc, 512)
            fm.quantize_ckv_per_token_head(q_buf, 512)

        t_f, t_u, t_q = time_graph(fused, n_rot), time_graph(unfused, n_rot), time_graph(quantisers_only, n_rot)
        nbytes = tokens * heads * (576 * 2 + 512 + 4 + 128) + tokens * (576 * 2 + 644)
        print("prologue tokens=%d heads=%d: fused %.1f us (%.0f GB/s algorithmic) | rope + 2 quantiser launches %.1f us | "
              "the 2 quantiser launches alone (no rope) %.1f us" % (tokens, heads, t_f, nbytes / t_f / 1e3, t_u, t_q),
              flush=True)


if __name__ == "__main__":
    main()
