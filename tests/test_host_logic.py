"""CPU: host-side logic of the Python mirrors (layout helpers, argument checking) -- no GPU needed."""
import os
import sys

import numpy as np
import types

import pytest
import torch

from helpers import quantize_lhs_1x128, quantize_rhs_128x128, similarity_diff
from sglang_fluentllm_b200 import deep_gemm as dg
from sglang_fluentllm_b200 import flash_mla_fp8 as fm
from sglang_fluentllm_b200 import fp8_kernel as fk


def test_col_major_tma_aligned_tensor_layout():
    x = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
    y = dg.get_col_major_tma_aligned_tensor(x)
    assert y.shape == (7, 3) and torch.equal(y, x)
    assert y.stride() == (1, 8)  # column-major, m padded to a multiple of 4 floats (16 bytes)
    assert dg.get_col_major_tma_aligned_tensor(y).data_ptr() == y.data_ptr()  # already aligned: no copy
    assert dg.get_tma_aligned_size(5, 4) == 8 and dg.get_tma_aligned_size(128, 4) == 128


def test_per_token_and_per_block_cast_roundtrip():
    torch.manual_seed(0)
    x = torch.randn(9, 256, dtype=torch.bfloat16)
    q, s = quantize_lhs_1x128(x)
    assert q.dtype == torch.float8_e4m3fn and s.shape == (9, 2)
    deq = q.float().view(9, 2, 128) * s.unsqueeze(-1)
    assert (deq.view(9, 256) - x.float()).abs().max() < 0.07 * x.float().abs().max()
    w = torch.randn(130, 256, dtype=torch.bfloat16)
    wq, ws = quantize_rhs_128x128(w)
    assert wq.shape == (130, 256) and ws.shape == (2, 2)


def test_scale_buffer_layout_matches_reference_contract():
    s = fk.create_per_token_group_quant_fp8_output_scale((33, 512), "cpu", 128, True, True)
    assert s.shape == (33, 4) and s.stride() == (1, 36)
    s = fk.create_per_token_group_quant_fp8_output_scale((33, 512), "cpu", 128, True, False)
    assert s.shape == (33, 4) and s.stride() == (1, 33)
    s = fk.create_per_token_group_quant_fp8_output_scale((2, 5, 256), "cpu", 128, False, False)
    assert s.shape == (2, 5, 2)


def test_gemm_argument_checks_match_reference():
    a = torch.zeros(4, 256).to(torch.float8_e4m3fn)
    a_s = torch.ones(4, 2)
    b = torch.zeros(64, 256).to(torch.float8_e4m3fn)
    b_s = torch.ones(1, 2)
    with pytest.raises(AssertionError):  # n % 64 (jit_kernels/gemm.py:169)
        dg.gemm_fp8_fp8_bf16_nt((a, a_s), (torch.zeros(32, 256).to(torch.float8_e4m3fn), b_s),
                                torch.zeros(4, 32, dtype=torch.bfloat16))
    with pytest.raises(AssertionError):  # out dtype
        dg.gemm_fp8_fp8_bf16_nt((a, a_s), (b, b_s), torch.zeros(4, 64, dtype=torch.float16))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        dg.gemm_fp8_fp8_bf16_nt((a, a_s), (b, b_s), torch.zeros(4, 64, dtype=torch.bfloat16))
    # m == 0 returns without touching the device (gemm.py:176-177)
    dg.gemm_fp8_fp8_bf16_nt((a[:0], a_s[:0]), (b, b_s), torch.zeros(0, 64, dtype=torch.bfloat16))


def test_mla_entry_points_reject_cpu_tensors():
    with pytest.raises(RuntimeError, match="CUDA"):
        fm.get_mla_metadata(torch.zeros(4, dtype=torch.int32), 16, 1)
    with pytest.raises(RuntimeError, match="CUDA"):
        fm.quantize_ckv_per_token_head(torch.zeros(2, 576, dtype=torch.bfloat16), 512)
    with pytest.raises(RuntimeError):
        fk.sglang_per_token_group_quant_fp8(torch.zeros(2, 256, dtype=torch.bfloat16), 128)


def test_fused_prologue_rejects_cpu_tensors_and_bad_caches():
    from sglang_fluentllm_b200 import rope
    q_nope = torch.zeros(2, 4, 512, dtype=torch.bfloat16)
    q_pe = torch.zeros(2, 4, 64, dtype=torch.bfloat16)
    k_pe = torch.zeros(2, 64, dtype=torch.bfloat16)
    cache = torch.zeros(16, 64)
    with pytest.raises(RuntimeError, match="CUDA"):
        rope.fused_rope_quantize_set_kv(torch.zeros(2, dtype=torch.int64), q_nope, q_pe, k_pe, cache, False)
    pool = types.SimpleNamespace(quant_method="none")
    with pytest.raises(ValueError, match="per_token_head"):
        rope.create_fused_set_kv_buffer_arg(q_nope, types.SimpleNamespace(layer_id=0),
                                            types.SimpleNamespace(token_to_kv_pool=pool, out_cache_loc=None))


def test_similarity_diff():
    x = torch.randn(100)
    assert similarity_diff(x, x) < 1e-12
    assert similarity_diff(x, -x) > 1.9


def test_fp8_config_and_linear_method_mirror_reference_rules():
    from sglang_fluentllm_b200 import fp8_linear as fl
    cfg = fl.Fp8Config.from_config({"quant_method": "fp8", "activation_scheme": "dynamic", "weight_block_size": [128, 128]})
    assert cfg.get_name() == "fp8" and cfg.is_checkpoint_fp8_serialized and cfg.weight_block_size == [128, 128]
    with pytest.raises(ValueError, match="Unsupported activation scheme"):
        fl.Fp8Config(activation_scheme="sometimes")
    with pytest.raises(ValueError, match="fp8-serialized"):
        fl.Fp8Config(is_checkpoint_fp8_serialized=False, weight_block_size=[128, 128])
    with pytest.raises(ValueError, match="dynamic activation scheme"):
        fl.Fp8Config(True, "static", None, [128, 128])
    cfg2 = fl.Fp8Config(True, "dynamic", ["lm_head"], [128, 128])
    assert cfg2.get_quant_method(torch.nn.Module(), prefix="lm_head") is None
    method = cfg2.get_quant_method(torch.nn.Module(), prefix="model.layers.0.self_attn.o_proj")
    assert isinstance(method, fl.Fp8LinearMethod) and method.block_quant


def test_fp8_linear_method_weights_and_checks():
    from sglang_fluentllm_b200 import fp8_linear as fl
    cfg = fl.Fp8Config(True, "dynamic", None, [128, 128])
    method = fl.Fp8LinearMethod(cfg, tp_size=2)
    layer = torch.nn.Module()
    method.create_weights(layer, 256, [384], 512, 384, torch.bfloat16)   # row parallel: K split over 2 ranks
    assert layer.weight.shape == (384, 256) and layer.weight.dtype == torch.float8_e4m3fn
    assert layer.weight_scale_inv.shape == (3, 2) and layer.input_scale is None
    with pytest.raises(ValueError, match="block_k"):
        method.create_weights(torch.nn.Module(), 192, [384], 384, 384, torch.bfloat16)
    with pytest.raises(ValueError, match="block_n"):
        method.create_weights(torch.nn.Module(), 256, [192, 64], 256, 256, torch.bfloat16)
    method.process_weights_after_loading(layer)
    # apply: non-bf16 output and unsupported shapes raise instead of falling back; CPU tensors never reach a kernel
    x = torch.zeros(4, 256, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no fallback"):
        method.apply(layer, x, output_dtype=torch.float16)
    with pytest.raises(RuntimeError):
        method.apply(layer, x)
    with pytest.raises(NotImplementedError):
        fl.Fp8LinearMethod(fl.Fp8Config())


def test_install_redirects_the_references_imports():
    """INTEGRATION.md section 2 as code: after install() the reference's `import flash_mla_fp8` / `import deep_gemm`
    resolve to the B200 mirrors; uninstall() restores the interpreter."""
    import importlib
    import sys
    from sglang_fluentllm_b200 import install as inst
    for name in ("flash_mla_fp8", "flash_mla_swap", "deep_gemm", "sglang.srt.layers.attention.flashmla_backend"):
        assert name not in sys.modules
    inst.install()
    try:
        from sglang_fluentllm_b200 import flashmla_backend as our_backend
        assert sys.modules["sglang.srt.layers.attention.flashmla_backend"] is our_backend   # base_spec_worker.py:233-240
        import deep_gemm as ref_name_dg
        import flash_mla_fp8 as ref_name_fm
        assert ref_name_fm is fm and ref_name_dg is dg
        swap = importlib.import_module("flash_mla_swap")
        assert swap.flash_mla_with_kvcache is fm.flash_mla_with_kvcache
        for attr in ("get_mla_metadata", "flash_mla_with_kvcache", "flash_mla_ckv_fp8_per_token",
                     "quantize_ckv_per_token_head", "dequantize_ckv_fused_indexed", "quantize_and_cache_k"):
            assert callable(getattr(ref_name_fm, attr))       # flash_mla_fp8/__init__.py:3-10
        for attr in ("gemm_fp8_fp8_bf16_nt", "m_grouped_gemm_fp8_fp8_bf16_nt_contiguous",
                     "m_grouped_gemm_fp8_fp8_bf16_nt_masked", "m_grouped_gemm_fp8_fp8_bf16_nt_offset",
                     "get_col_major_tma_aligned_tensor", "get_m_alignment_for_contiguous_layout"):
            assert callable(getattr(ref_name_dg, attr))
        sys.modules["deep_gemm"] = types.ModuleType("deep_gemm")   # a foreign module in the way
        with pytest.raises(RuntimeError, match="already imported"):
            inst.install()
        inst.install(force=True)
        assert sys.modules["deep_gemm"] is dg
    finally:
        inst.uninstall()
        sys.modules.pop("deep_gemm", None)
    assert "flash_mla_fp8" not in sys.modules and "flash_mla_swap" not in sys.modules
    assert "sglang.srt.layers.attention.flashmla_backend" not in sys.modules
    fake = types.ModuleType("attn_initializer")
    inst.patch_attention_initializer(fake)
    from sglang_fluentllm_b200 import flashmla_backend as fb
    assert fake.FlashMLABackend is fb.FlashMLABackend
    assert fake.FlashMLAMultiStepDecodeBackend is fb.FlashMLAMultiStepDecodeBackend


REFERENCE_PY = "/root/reference/python"


@pytest.mark.skipif(not os.path.isdir(REFERENCE_PY), reason="the reference tree only exists in the build container")
def test_backend_metadata_path_on_the_references_own_pool_and_allocator(monkeypatch):
    """The plug-in reads ``model_runner.kv_allocator.req_to_page`` and ``forward_batch.token_to_kv_pool`` -- here the
    REFERENCE's KVAllocator (mem_cache/allocator.py) and MLATokenToKVPool (mem_cache/memory_pool.py, per-token FP8
    layout), imported from /root/reference through the stub hook of tests/golden/make_golden.py and living on the CPU.
    init_forward_metadata / the CUDA-graph capture + replay protocol must produce the allocator's block table bit for bit
    and hand the kernel module the lengths the reference hands it (flashmla_backend.py:258-405); the metadata kernel
    itself is replaced by the oracle (no GPU here), so the integers are the reference scheduler's."""
    import importlib.util
    from types import SimpleNamespace
    from oracle import mla as omla
    from sglang_fluentllm_b200 import flashmla_backend as fb
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    finder = mg._Finder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REFERENCE_PY)
    before = set(sys.modules)
    try:
        from sglang.srt.mem_cache.allocator import KVAllocator
        from sglang.srt.mem_cache.memory_pool import MLATokenToKVPool
        alloc = KVAllocator(size=64 * 64, device="cpu", max_batch_size=8, max_context_len=1024, page_size=64)
        pool = MLATokenToKVPool(size=64 * 64, model_dtype=torch.bfloat16, dtype=torch.float8_e4m3fn,
                                quant_method="per_token_head", kv_lora_rank=512, qk_rope_head_dim=64, layer_num=2,
                                device="cpu", enable_memory_saver=False, max_batch_size=8, max_context_len=1024,
                                page_size=64, rank=0, enable_alt_stream=False)
        ctx = {3: 200, 1: 65, 7: 1000}
        for req, n in ctx.items():
            assert alloc.alloc(req, n, 0) is not None
        seen = []

        def oracle_metadata(cache_seqlens, num_heads_per_head_k, num_heads_k, fixed_overhead_blocks=5, num_sm_parts=None):
            seen.append((cache_seqlens.clone(), num_heads_per_head_k, num_heads_k))
            assert num_sm_parts is None  # 16 q rows: the reference's rule
            meta, ns = omla.get_mla_metadata(cache_seqlens.numpy(), num_heads_per_head_k, num_heads_k, 148,
                                             fixed_overhead_blocks=fixed_overhead_blocks)
            return torch.from_numpy(meta), torch.from_numpy(ns)

        monkeypatch.setattr(fb.flash_mla_fp8, "get_mla_metadata", oracle_metadata)
        cfg = SimpleNamespace(scaling=192 ** -0.5, num_attention_heads=128, kv_lora_rank=512, qk_nope_head_dim=128,
                              qk_rope_head_dim=64, v_head_dim=512, context_len=1024)
        mr = SimpleNamespace(model_config=cfg, server_args=SimpleNamespace(speculative_algorithm=None),
                             kv_cache_dtype=pool.dtype, kv_cache_quant_method=pool.quant_method, kv_allocator=alloc,
                             dtype=torch.bfloat16, device="cpu")
        fb.set_attention_tp_size_getter(lambda: 8)
        try:
            backend = fb.FlashMLABackend(mr)
        finally:
            fb.set_attention_tp_size_getter(None)
        assert backend.num_q_heads == 16
        reqs = torch.tensor([7, 1, 3])
        seq_lens = torch.tensor([ctx[7] + 1, ctx[1] + 1, ctx[3] + 1])
        fbatch = SimpleNamespace(batch_size=3, seq_lens=seq_lens, req_pool_indices=reqs, token_to_kv_pool=pool,
                                 forward_mode=fb.ForwardMode.DECODE)
        backend.init_forward_metadata(fbatch)
        md = backend.forward_metadata
        assert torch.equal(md.block_table, alloc.req_to_page[reqs]) and md.block_table.dtype == torch.int32
        lens, m, hk = seen[-1]
        assert lens.dtype == torch.int32 and lens.tolist() == seq_lens.tolist() and (m, hk) == (16, 1)
        # (the backend asks for the schedule with the per-request cost tuned for the B200 kernels)
        want_meta, want_ns = omla.get_mla_metadata(seq_lens.numpy(), 16, 1, 148,
                                                   fixed_overhead_blocks=fm.tuned_overhead_blocks(16))
        assert np.array_equal(md.flashmla_metadata.numpy(), want_meta) and np.array_equal(md.num_splits.numpy(), want_ns)
        # the three tensors per layer the kernel call receives are the reference pool's own (views, no copies)
        k_lora, k_scale, k_rope = pool.get_key_buffer(1)
        assert k_lora.shape[1:] == (1, 512) and k_scale.shape[1:] == (1, 1) and k_rope.shape[1:] == (1, 64)
        assert k_lora.view(-1, 64, 1, 512).data_ptr() == k_lora.data_ptr()
        # CUDA-graph protocol (cuda_graph_runner.py:405-520): static buffers, capture, replay with grown sequences
        backend.init_cuda_graph_state(max_bs=4)
        assert backend.cuda_graph_kv_indices.shape == (4, (1024 + 63 + 256) // 64) and int(backend.cuda_graph_kv_indices.min()) == 1
        backend.init_forward_metadata_capture_cuda_graph(3, 3, reqs, seq_lens, fb.ForwardMode.DECODE, None)
        static_bt = backend.forward_metadata.block_table
        assert static_bt.data_ptr() == backend.cuda_graph_kv_indices.data_ptr()
        assert torch.equal(static_bt[:, : alloc.req_to_page.shape[1]], alloc.req_to_page[reqs])
        assert alloc.alloc(1, 70, ctx[1]) is not None                 # request 1 grows past a page boundary
        seq_lens2 = seq_lens + torch.tensor([0, 70, 0])
        backend.init_forward_metadata_replay_cuda_graph(3, reqs, seq_lens2, int(seq_lens2.sum()), fb.ForwardMode.DECODE, None,
                                                        seq_lens_cpu=seq_lens2)
        assert backend.forward_metadata.block_table.data_ptr() == static_bt.data_ptr()   # refreshed in place
        assert torch.equal(backend.forward_metadata.block_table[:, : alloc.req_to_page.shape[1]], alloc.req_to_page[reqs])
        assert seen[-1][0].tolist() == seq_lens2.tolist()
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REFERENCE_PY)
        for name in set(sys.modules) - before:
            if name.split(".")[0] in mg.STUBS or name.split(".")[0] == "sglang":
                del sys.modules[name]
