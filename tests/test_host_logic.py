"""CPU: host-side logic of the Python mirrors (layout helpers, argument checking) -- no GPU needed."""
import numpy as np
import types

import pytest
import torch

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
    q, s = dg.per_token_cast_to_fp8(x)
    assert q.dtype == torch.float8_e4m3fn and s.shape == (9, 2)
    deq = q.float().view(9, 2, 128) * s.unsqueeze(-1)
    assert (deq.view(9, 256) - x.float()).abs().max() < 0.07 * x.float().abs().max()
    w = torch.randn(130, 256, dtype=torch.bfloat16)
    wq, ws = dg.per_block_cast_to_fp8(w)
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


def test_calc_diff():
    x = torch.randn(100)
    assert dg.calc_diff(x, x) < 1e-12
    assert dg.calc_diff(x, -x) > 1.9


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
    for name in ("flash_mla_fp8", "flash_mla_swap", "deep_gemm"):
        assert name not in sys.modules
    inst.install()
    try:
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
    fake = types.ModuleType("attn_initializer")
    inst.patch_attention_initializer(fake)
    from sglang_fluentllm_b200 import flashmla_backend as fb
    assert fake.FlashMLABackend is fb.FlashMLABackend
    assert fake.FlashMLAMultiStepDecodeBackend is fb.FlashMLAMultiStepDecodeBackend
