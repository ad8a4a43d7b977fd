"""Generate the golden fixtures under tests/golden/ from the REFERENCE's own Python, imported from /root/reference.

Run once in the build container (`python tests/golden/make_golden.py`); the .npz files are committed, the reference
tree is never needed at test time (it does not exist on the GPU box).

What can be pinned this way:
  * MLATokenToKVPool per_token_head torch fall-back quantiser + de-quantiser
    (python/sglang/srt/mem_cache/memory_pool.py:826-831, 872-880), via the real class on device="cpu";
  * KVAllocator page-table bookkeeping (python/sglang/srt/mem_cache/allocator.py:41-102);
  * TorchNativeAttnBackend._run_sdpa_forward_decode (layers/attention/torch_native_backend.py:275-343);
  * native_per_token_group_quant_fp8 / native_w8a8_block_fp8_matmul (python/sglang/test/test_block_fp8.py:15-40,
    89-141) -- the file itself does not import (stale names), so the two functions are extracted by AST;
  * per_token_cast_to_fp8 / per_block_cast_to_fp8 (3rdparty/deep_gemm/tests/test_core.py:9-25), same extraction.
  * DeepseekScalingRotaryEmbedding.forward_native (python/sglang/srt/layers/rotary_embedding.py:804-846), both pairings
    (`python tests/golden/make_golden.py rope` regenerates only this one);
  * the reference's own attention test up to the kernel call: inputs, test-side quantisation and the fp32 `ref_mla`
    outputs its kernel is compared with (3rdparty/flashmla-fp8/tests/test_flash_mla_ckv_fp8_per_token.py:12-101);
  * the reference's own quantiser test up to the kernel calls: generator, torch reference outputs
    (3rdparty/flashmla-fp8/tests/test_cache_kernels.py:11-103);
  * MLATokenToKVPool.get_contiguous_buf_infos / get_layerwise_buf_info_offsets / move_kv_cache
    (mem_cache/memory_pool.py:747-806), both pool layouts.
  * the reference's metadata / quantiser / de-quantiser CUDA kernels themselves, compiled for the host from their sources
    and run on CPU threads (oracle/ref_build/build_ref.py -> oracle/_ref/libflashmla_ref_cpu.so): ref_kernels_cpu.npz.
The sm_90a attention kernel (wgmma, TMA, CuTe) cannot run here: no fixture of its own outputs.
"""
import ast
import importlib.abc
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
STUBS = {"eps", "orjson", "flash_mla_fp8", "flash_mla_swap", "flash_mla", "deep_gemm", "deep_gemm_oss", "deep_ep",
         "fast_hadamard_transform", "torchao", "outlines", "decord"}


class _Stub(types.ModuleType):
    __path__ = []

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return None


class _Finder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in STUBS:
            return importlib.machinery.ModuleSpec(name, self, is_package=True)
        return None

    def create_module(self, spec):
        return _Stub(spec.name)

    def exec_module(self, module):
        pass


def extract_functions(path, names):
    src = open(path).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "Tuple": tuple}
    from typing import Tuple
    ns["Tuple"] = Tuple
    ns["ceil_div"] = lambda a, b: (a + b - 1) // b
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def make_pool_disagg():
    """PD-disaggregation descriptors and slot moves of the per_token_head pool through the reference's own class on CPU
    (python/sglang/srt/mem_cache/memory_pool.py:747-806)."""
    from sglang.srt.mem_cache.memory_pool import MLATokenToKVPool
    torch.manual_seed(2)
    out = {}
    for qm, dtype in (("per_token_head", torch.float8_e4m3fn), ("none", torch.bfloat16)):
        pool = MLATokenToKVPool(size=64, model_dtype=torch.bfloat16, dtype=dtype, quant_method=qm, kv_lora_rank=512,
                                qk_rope_head_dim=64, layer_num=3, device="cpu", enable_memory_saver=False,
                                max_batch_size=4, max_context_len=64, page_size=64, rank=0, enable_alt_stream=False)
        ptrs, lens, item_lens = pool.get_contiguous_buf_infos()
        bufs = [t for kv in pool.kv_buffer for t in (kv if isinstance(kv, tuple) else (kv,))]
        order = [[b.data_ptr() for b in bufs].index(p) for p in ptrs]   # which tensor (layer-major flat index) each is
        out["lens_" + qm], out["item_lens_" + qm], out["order_" + qm] = np.array(lens), np.array(item_lens), np.array(order)
        out["offsets_" + qm] = np.array(pool.get_layerwise_buf_info_offsets(5))
        try:
            out["size_bytes_" + qm] = np.array(pool.get_kv_size_bytes())
        except AttributeError:   # :779-784 calls get_tensor_size_bytes on the per-layer tuple of the FP8 layout and raises
            out["size_bytes_" + qm] = np.array(-1)
        if qm == "per_token_head":
            for kv in pool.kv_buffer:
                kv[0].copy_(torch.randint(0, 255, kv[0].shape, dtype=torch.uint8).view(kv[0].dtype))
                kv[1].copy_(torch.rand(kv[1].shape))
                kv[2].copy_(torch.randn(kv[2].shape).to(torch.bfloat16))
            before = [[t.clone() for t in kv] for kv in pool.kv_buffer]
            src, tgt = torch.tensor([64, 65, 70, 100]), torch.tensor([90, 91, 64, 127])
            pool._kv_copy_config = None   # the native path (:748-764); the Triton tiled copy needs a GPU
            pool.move_kv_cache(tgt, src)
            out["move_src"], out["move_tgt"] = src.numpy(), tgt.numpy()
            for l, kv in enumerate(pool.kv_buffer):
                for i, name in enumerate(("lora", "scale", "rope")):
                    out["before_%s_%d" % (name, l)] = before[l][i].view(torch.uint8).numpy() if i == 0 else before[l][i].float().numpy()
                    out["after_%s_%d" % (name, l)] = kv[i].view(torch.uint8).numpy() if i == 0 else kv[i].float().numpy()
    np.savez_compressed(os.path.join(HERE, "pool_disagg.npz"), **out)


def make_mla_test_reference():
    """The reference's own attention test re-run on CPU up to the kernel call
    (3rdparty/flashmla-fp8/tests/test_flash_mla_ckv_fp8_per_token.py): its input construction (:48-69, NaN-poisoned
    padding), its fp32 reference `ref_mla` / `scaled_dot_product_attention` (:12-30, :75-92, extracted by AST) and its
    test-side per-token quantisation of q and of the paged K (:94-101).  The expected outputs are what the reference
    compares its kernel against with relative_L1 < 5e-2 (:33-40)."""
    import math
    import random
    path = os.path.join(REF, "3rdparty/flashmla-fp8/tests/test_flash_mla_ckv_fp8_per_token.py")
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "math": math}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "scaled_dot_product_attention":
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    sdpa = ns["scaled_dot_product_attention"]
    torch.manual_seed(0)
    random.seed(0)
    old_dtype = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)   # main(torch.bfloat16), :141-146
    out = {}
    try:
        for tag, (b, s_q, mean_sk, h_q, causal, varlen) in {"decode": (3, 1, 120, 16, False, True),
                                                             "verify": (2, 2, 100, 8, True, True)}.items():
            h_kv, d, dv, block_size = 1, 576, 512, 64
            cache_seqlens = torch.full((b,), mean_sk, dtype=torch.int32)
            if varlen:
                for i in range(b):
                    cache_seqlens[i] = max(random.normalvariate(mean_sk, mean_sk / 2), s_q)
            max_seqlen = cache_seqlens.max().item()
            max_seqlen_pad = -(-max_seqlen // (block_size * 4)) * block_size * 4
            q = torch.randn(b, s_q, h_q, d)
            block_table = torch.arange(b * max_seqlen_pad // block_size, dtype=torch.int32).view(b, max_seqlen_pad // block_size)
            blocked_k = torch.randn(block_table.numel(), block_size, h_kv, d)
            for i in range(b):
                blocked_k.view(b, max_seqlen_pad, h_kv, d)[i, cache_seqlens[i].item():] = float("nan")
            blocked_v = blocked_k[..., :dv]
            o_ref = torch.empty(b, s_q, h_q, dv, dtype=torch.float32)
            lse_ref = torch.empty(b, h_q, s_q, dtype=torch.float32)
            for i in range(b):
                begin = i * max_seqlen_pad
                end = begin + cache_seqlens[i]
                O, LSE = sdpa(q[i].transpose(0, 1), blocked_k.view(-1, h_kv, d)[begin:end].transpose(0, 1),
                              blocked_v.view(-1, h_kv, dv)[begin:end].transpose(0, 1), h_q=h_q, h_kv=h_kv, is_causal=causal)
                o_ref[i] = O.transpose(0, 1)
                lse_ref[i] = LSE
            q_nope = q[..., :dv]
            q_scale = q_nope.float().abs().amax(dim=-1, keepdim=True) / 448.0
            q_rope = (q[..., dv:].float() / q_scale).to(q.dtype)
            q_nope_fp8 = (q_nope.float() / q_scale).to(torch.float8_e4m3fn)
            k_lora = blocked_k[..., :dv]
            k_scale = k_lora.float().abs().amax(dim=-1, keepdim=True) / 448.0
            k_rope = (blocked_k[..., dv:].float() / k_scale).to(q.dtype)
            k_lora_fp8 = (k_lora.float() / k_scale).to(torch.float8_e4m3fn)
            # bf16 tensors are stored as their bit patterns (uint16) to keep the fixture small
            for name, t in (("q_bf16bits", q.view(torch.int16)), ("blocked_k_bf16bits", blocked_k.view(torch.int16)),
                            ("cache_seqlens", cache_seqlens),
                            ("block_table", block_table), ("out", o_ref), ("lse", lse_ref),
                            # of the test-side quantisation only the scales are stored (the e4m3 / bf16 tensors follow from
                            # them by the three lines at :94-101, which the consuming test restates); a checksum pins the codes
                            ("q_scale", q_scale), ("k_scale", k_scale),
                            ("q_nope_fp8_sum", q_nope_fp8.view(torch.uint8).long().sum()),
                            ("k_lora_fp8_sum", k_lora_fp8.view(torch.uint8).long().sum())):
                out["%s_%s" % (tag, name)] = t.numpy()
            out["%s_causal" % tag] = np.array(causal)
    finally:
        torch.set_default_dtype(old_dtype)
    np.savez_compressed(os.path.join(HERE, "mla_test_reference.npz"), **out)


def make_cache_kernel_test_reference():
    """The reference's own quantiser test re-run on CPU up to the kernel calls
    (3rdparty/flashmla-fp8/tests/test_cache_kernels.py): its input generator (:11-57, extracted by AST; rows with
    different ranges and offsets) and its torch reference `ref_quant_and_cache` (:88-103).  The reference demands
    relative_L1 < 1e-3 between these and its CUDA quantisers for all six outputs (:59-67, :141-146)."""
    path = os.path.join(REF, "3rdparty/flashmla-fp8/tests/test_cache_kernels.py")
    (gen,) = extract_functions(path, ["generate_random_tensor_with_varying_ranges"])
    old_dtype = torch.get_default_dtype()
    torch.set_default_dtype(torch.bfloat16)   # main(torch.bfloat16), :153-158
    try:
        b, s_q, h_q, d, dv = 4, 2, 16, 576, 512
        block_num = b * s_q * 4
        query = gen(num_vectors=b * s_q * h_q, vector_length=d).reshape(b, s_q, h_q, d)
        key = gen(num_vectors=b * s_q, vector_length=d).reshape(b, s_q, 1, d)
        assert query.dtype == torch.bfloat16
        loc = torch.randperm(block_num * 64, dtype=torch.int32)[: b * s_q]
        q_nope, q_rope = query[..., :dv].clone(), query[..., dv:].clone()
        q_scale = q_nope.float().abs().amax(dim=-1, keepdim=True) / 448.0
        q_nope = (q_nope.float() / q_scale).to(torch.float8_e4m3fn).view(torch.uint8)
        q_rope = (q_rope.float() / q_scale).to(query.dtype)
        k_lora, k_rope = key[..., :dv], key[..., dv:]
        k_scale = k_lora.float().abs().amax(dim=-1, keepdim=True) / 448.0
        k_lora = (k_lora.float() / k_scale).to(torch.float8_e4m3fn).view(torch.uint8)
        k_rope = (k_rope.float() / k_scale).to(query.dtype)
        np.savez_compressed(
            os.path.join(HERE, "cache_kernel_test_reference.npz"),
            query_bf16bits=query.view(torch.int16).numpy(), key_bf16bits=key.view(torch.int16).numpy(), loc=loc.numpy(),
            q_nope=q_nope.numpy(), q_scale=q_scale.float().numpy(), q_rope=q_rope.float().numpy(),
            k_lora=k_lora.reshape(-1, 1, dv).numpy(), k_scale=k_scale.float().reshape(-1, 1, 1).numpy(),
            k_rope=k_rope.float().reshape(-1, 1, d - dv).numpy(), num_slots=np.array(block_num * 64))
    finally:
        torch.set_default_dtype(old_dtype)


def make_ref_kernels_cpu():
    """Outputs of the reference's own metadata / quantiser / de-quantiser / activation-quantiser kernels, compiled for the host from their
    sources by oracle/ref_build/build_ref.py and executed on CPU threads (oracle/ref_cpu.py).  The fixture lets the
    bit-exact pin run where oracle/_ref is absent."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    from oracle.ref_build import build_ref
    build_ref.build(REF)
    from oracle import ref_cpu
    rng = np.random.default_rng(2024)
    out = {}
    cases = [(148, 16, 1), (148, 16, 128), (148, 128, 128), (148, 64, 37), (132, 64, 200), (74, 128, 5), (8, 16, 3)]
    for i, (sm, hpk, b) in enumerate(cases):
        seqlens = rng.integers(1, 9000, size=b).astype(np.int32) if i % 2 else np.full(b, 4096, np.int32)
        meta, ns = ref_cpu.get_mla_metadata(seqlens, hpk, 1, sm)
        out["meta_case_%d" % i] = np.array([sm, hpk, b])
        out["meta_seqlens_%d" % i], out["meta_out_%d" % i], out["meta_num_splits_%d" % i] = seqlens, meta, ns
    out["meta_cases"] = np.array(len(cases))
    f32 = (rng.standard_normal((48, 4, 576)) * rng.uniform(0.001, 40, (48, 4, 1))).astype(np.float32)
    f32[3] = 0                       # all-zero rows: scale = 1e-12
    f32[4, 0, :512] = 7.0            # constant row: every element on the e4m3 maximum
    q_bits = torch.from_numpy(f32).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    out["q_bits"] = q_bits
    out["q_nope"], out["q_scale"], out["q_rope_bits"] = ref_cpu.quantize_ckv_per_token_head(q_bits)
    slots, n = 64 * 5, 40
    key = (rng.standard_normal((n, 1, 576)) * rng.uniform(0.001, 40, (n, 1, 1))).astype(np.float32)
    k_bits = torch.from_numpy(key).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    loc = rng.permutation(slots)[:n].astype(np.int32)
    kl, ks, kr = ref_cpu.pool_arrays(slots)
    ref_cpu.quantize_and_cache_k(k_bits, kl, ks, kr, loc)
    idx = rng.choice(loc, size=29).astype(np.int32)
    lo, ro = ref_cpu.dequantize_ckv_fused_indexed(kl, kr, ks, idx)
    lo2, ro2 = ref_cpu.dequantize_ckv_fused_indexed(kl, kr, ks, idx, max_num_tokens=6)
    assert np.array_equal(lo, lo2) and np.array_equal(ro, ro2)
    out.update(k_bits=k_bits, k_loc=loc, k_slots=np.array(slots), k_lora=kl.copy(), k_scale=ks.copy(), k_rope_bits=kr.copy(),
               deq_idx=idx, deq_lora_bits=lo, deq_rope_bits=ro)
    # G-2: the 1x128 activation quantiser kernel (3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-437)
    xa = (rng.standard_normal((37, 1536)) * rng.uniform(0.001, 50, (37, 1))).astype(np.float32)
    xa[0, :128] = 0                 # an all-zero group: scale = eps / 448
    a_bits = torch.from_numpy(xa).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    aq, a_s = ref_cpu.per_token_group_quant_fp8(a_bits)
    out.update(act_bits=a_bits, act_q=aq, act_scales=a_s)
    # a-5: the split-KV combine kernel (ckv_fp8_per_token/mla_combine.cu:16-154) on fp32 partials in the layout the
    # attention kernel leaves them in (o_accum [splits, M, 512], lse_accum [splits, M] in the log2 domain)
    cb, cM, parts = 3, 4, 8
    c_ns = np.array([0, 2, 3, 8], np.int32)                       # 2, 1 and 5 splits
    c_oacc = rng.standard_normal((cb + parts, cM, 512)).astype(np.float32)
    c_lacc = (rng.standard_normal((cb + parts, cM)) * 3 + 5).astype(np.float32)
    c_lacc[3, 1] = -np.inf                                         # a split that saw no key for this row
    c_lacc[4:8, 2] = -np.inf                                       # all but one
    c_out, c_lse = ref_cpu.mla_combine(c_oacc, c_lacc, c_ns, cb, 1, cM, parts)
    out.update(comb_o_accum=c_oacc, comb_lse_accum=c_lacc, comb_num_splits=c_ns, comb_out_bits=c_out, comb_lse=c_lse)
    np.savez_compressed(os.path.join(HERE, "ref_kernels_cpu.npz"), **out)


def make_rope():
    """Rotary embedding of the MLA rope dims through the reference's own classes on CPU
    (python/sglang/srt/layers/rotary_embedding.py: DeepseekScalingRotaryEmbedding.forward_native :804-846 with the
    DeepSeek-V3 yarn parameters, interleaved pairs; the same class with is_neox_style=True for the LongCat pairing)."""
    from sglang.srt.layers.rotary_embedding import DeepseekScalingRotaryEmbedding
    torch.manual_seed(1)
    tokens, heads = 24, 5
    positions = torch.cat([torch.tensor([0, 1, 2, 63, 64, 4095, 4096, 16383]), torch.randint(0, 16384, (tokens - 8,))])
    q_pe = (torch.randn(tokens, heads, 64) * 3).to(torch.bfloat16)
    k_pe = (torch.randn(tokens, 1, 64) * 3).to(torch.bfloat16)
    out = {"positions": positions.numpy(), "q_pe": q_pe.float().numpy(), "k_pe": k_pe.float().numpy()}
    for name, neox in (("gptj", False), ("neox", True)):
        rope = DeepseekScalingRotaryEmbedding(64, 64, 4096, 10000, neox, 40.0, torch.float32, beta_fast=32, beta_slow=1,
                                              mscale=1.0, mscale_all_dim=1.0, device="cpu")
        cache = rope.cos_sin_cache.float()
        q_out, k_out = rope.forward_native(positions, q_pe.clone(), k_pe.clone())
        assert q_out.dtype == torch.bfloat16
        out["cos_sin_rows_" + name] = cache[positions].numpy()   # the rows the vectors use (the full table is 40 MB)
        out["cache_shape_" + name] = np.array(cache.shape)
        # the NeoX branch assumes [batch, seq] positions and returns a leading 1 (rotary_embedding.py:826-830)
        out["q_out_" + name] = q_out.float().numpy().reshape(q_pe.shape)
        out["k_out_" + name] = k_out.float().numpy().reshape(k_pe.shape)
    np.savez_compressed(os.path.join(HERE, "rope.npz"), **out)


def main():
    sys.meta_path.insert(0, _Finder())
    sys.path.insert(0, os.path.join(REF, "python"))
    only = {"rope": make_rope, "pool_disagg": make_pool_disagg, "mla_test_reference": make_mla_test_reference,
            "cache_kernel_test_reference": make_cache_kernel_test_reference, "ref_kernels_cpu": make_ref_kernels_cpu}
    if sys.argv[1:] and set(sys.argv[1:]) <= set(only):  # regenerate only these fixtures
        for name in sys.argv[1:]:
            only[name]()
        return
    torch.manual_seed(0)

    # ---- pool fall-back quantiser / de-quantiser ----
    from sglang.srt.mem_cache.memory_pool import MLATokenToKVPool
    pool = MLATokenToKVPool(size=512, model_dtype=torch.bfloat16, dtype=torch.float8_e4m3fn,
                            quant_method="per_token_head", kv_lora_rank=512, qk_rope_head_dim=64, layer_num=1,
                            device="cpu", enable_memory_saver=False, max_batch_size=4, max_context_len=256,
                            page_size=64, rank=0, enable_alt_stream=False)
    layer = types.SimpleNamespace(layer_id=0)
    key = (torch.randn(96, 1, 576) * torch.rand(96, 1, 1) * 4).to(torch.bfloat16)
    key[5] = 0  # all-zero row exercises the clamp
    loc = torch.randperm(512)[:96] + 64
    for buf in pool.kv_buffer[0]:
        buf.zero_()
    pool.set_kv_buffer(layer, loc, key, None)
    k_lora, k_scale, k_rope = pool.kv_buffer[0]
    deq_lora, deq_rope = pool.get_key_split_contiguous(0, loc)
    np.savez_compressed(
        os.path.join(HERE, "pool_per_token_head.npz"),
        key=key.float().numpy(), loc=loc.numpy().astype(np.int32),
        k_lora=k_lora.view(torch.uint8).numpy(), k_scale=k_scale.numpy(), k_rope=k_rope.float().numpy(),
        deq_lora=deq_lora.float().numpy(), deq_rope=deq_rope.float().numpy())

    # ---- allocator ----
    from sglang.srt.mem_cache.allocator import KVAllocator
    alloc = KVAllocator(size=64 * 40, device="cpu", max_batch_size=4, max_context_len=640, page_size=64)
    trace = []
    lens = [0, 0, 0]
    for req, need in [(0, 130), (1, 64), (2, 1), (0, 1), (1, 1), (2, 70), (0, 62), (1, 200)]:
        kv_loc = alloc.alloc(req, need, lens[req])
        lens[req] += need
        trace.append((req, need, kv_loc.numpy().astype(np.int32)))
    np.savez_compressed(
        os.path.join(HERE, "allocator.npz"),
        reqs=np.array([t[0] for t in trace], dtype=np.int32), needs=np.array([t[1] for t in trace], dtype=np.int32),
        kv_locs=np.concatenate([t[2] for t in trace]), req_to_page=alloc.req_to_page.numpy().astype(np.int32),
        final_lens=np.array(lens, dtype=np.int32))

    # ---- torch-native SDPA decode ----
    from sglang.srt.layers.attention.torch_native_backend import TorchNativeAttnBackend
    b, h, total = 3, 16, 700
    seq_lens = torch.tensor([130, 64, 257])
    k_cache = torch.randn(total + 64, 1, 576).to(torch.bfloat16)
    v_cache = k_cache[..., :512]
    req_to_token = torch.zeros(4, 300, dtype=torch.int64)
    perm = torch.randperm(total) + 64
    off = 0
    for i, n in enumerate(seq_lens.tolist()):
        req_to_token[i, :n] = perm[off:off + n]
        off += n
    q = torch.randn(b, h, 576).to(torch.bfloat16)
    out = torch.zeros(b, h, 512, dtype=torch.bfloat16)
    backend = TorchNativeAttnBackend.__new__(TorchNativeAttnBackend)
    backend._run_sdpa_forward_decode(q, out, k_cache, v_cache, req_to_token, torch.arange(b), seq_lens,
                                     scaling=576 ** -0.5, enable_gqa=True, causal=False)
    np.savez_compressed(os.path.join(HERE, "native_sdpa_decode.npz"), q=q.float().numpy(),
                        k_cache=k_cache.float().numpy(), req_to_token=req_to_token.numpy(), seq_lens=seq_lens.numpy(),
                        out=out.float().numpy())

    # ---- block fp8 oracles ----
    quant, matmul = extract_functions(os.path.join(REF, "python/sglang/test/test_block_fp8.py"),
                                      ["native_per_token_group_quant_fp8", "native_w8a8_block_fp8_matmul"])
    x = (torch.randn(37, 512) * 3).to(torch.bfloat16)
    x[3, 128:256] = 0
    xq, xs = quant(x, 128)
    w = torch.randn(192, 512).to(torch.bfloat16)
    cast_tok, cast_blk = extract_functions(os.path.join(REF, "3rdparty/deep_gemm/tests/test_core.py"),
                                           ["per_token_cast_to_fp8", "per_block_cast_to_fp8"])
    wq, ws = cast_blk(w)
    c = matmul(xq, wq, xs, ws, [128, 128], output_dtype=torch.float32)
    xq2, xs2 = cast_tok(x)
    np.savez_compressed(os.path.join(HERE, "block_fp8.npz"), x=x.float().numpy(), xq=xq.view(torch.uint8).numpy(),
                        xs=xs.numpy(), w=w.float().numpy(), wq=wq.view(torch.uint8).numpy(), ws=ws.numpy(),
                        c=c.numpy(), xq_cast=xq2.view(torch.uint8).numpy(), xs_cast=xs2.numpy())
    make_rope()
    make_pool_disagg()
    make_mla_test_reference()
    make_cache_kernel_test_reference()
    make_ref_kernels_cpu()
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
