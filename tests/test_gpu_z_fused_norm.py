"""f-2: fused q_a / kv_a RMSNorm with the 1x128 quantiser of the q_b_proj input fused in (b200_qk_rmsnorm_quant behind
sglang_fluentllm_b200/layernorm.py::FusedRMSNorm) against the oracle and against the unfused sequence."""
import numpy as np
import pytest
import torch

from oracle import gemm as ogemm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ulp_ok(got, want):
    """bf16 values (as float32): equal or adjacent."""
    g, w = got.view(np.uint32) >> 16, want.view(np.uint32) >> 16
    return np.abs(g.astype(np.int64) - w.astype(np.int64)) <= 1


@pytest.mark.parametrize("m,q_dim", [(1, 1536), (37, 1536), (128, 1536), (5, 512)])
def test_fused_qk_rmsnorm_and_block_quant(m, q_dim):
    from sglang_fluentllm_b200 import fp8_kernel as fk
    from sglang_fluentllm_b200 import layernorm as ln
    torch.manual_seed(m)
    width = q_dim + 512 + 64                        # q_a | kv_a | k_pe, the fused qkv_a projection output
    qkv = (torch.randn(m, width, device=DEV) * 3).to(torch.bfloat16)
    qkv[0, :128] = 0                                # an all-zero quantiser group
    norm = ln.FusedRMSNorm(ln.RMSNorm(q_dim, 1e-6, DEV), ln.RMSNorm(512, 1e-6, DEV))
    norm.q_a_norm.weight.data.copy_((torch.rand(q_dim, device=DEV) + 0.5).to(torch.bfloat16))
    norm.kv_a_norm.weight.data.copy_((torch.rand(512, device=DEV) + 0.5).to(torch.bfloat16))
    want_q, want_kv = ogemm.fused_qk_rmsnorm(qkv.float().cpu().numpy(), q_dim, 512, norm.weight_q_a.float().cpu().numpy(),
                                             norm.weight_kv_a.float().cpu().numpy(), 1e-6)
    rope_before = qkv[:, q_dim + 512:].clone()
    # (1) the reference's call shape: q_a -> contiguous buffer, kv_a in place (deepseek_v2.py:812-820)
    a = qkv.clone()
    q, latent = a.split([q_dim, 512 + 64], dim=-1)
    q_contiguous = torch.empty_like(q)
    norm(input_q_a=q, input_kv_a=latent[..., :512], output_q_a=q_contiguous)
    assert q_contiguous.is_contiguous()
    assert _ulp_ok(q_contiguous.float().cpu().numpy(), want_q).all()
    assert _ulp_ok(a[:, q_dim:q_dim + 512].float().cpu().numpy(), want_kv).all()
    assert torch.equal(a[:, q_dim + 512:], rope_before) and torch.equal(a[:, :q_dim], qkv[:, :q_dim])  # untouched columns
    # (2) fused block quantisation: bit-identical to quantising (1)'s bf16 output in a second launch
    b = qkv.clone()
    q8, scales = norm.forward_block_quant(b, q_dim)
    ref8, ref_s = fk.sglang_per_token_group_quant_fp8(q_contiguous, 128, column_major_scales=True, scale_tma_aligned=True)
    assert torch.equal(q8.view(torch.uint8), ref8.view(torch.uint8))
    assert torch.equal(scales, ref_s) and scales.stride() == ref_s.stride()
    assert torch.equal(b[:, q_dim:q_dim + 512], a[:, q_dim:q_dim + 512])
    # and against the oracle quantiser on the kernel's own bf16 output
    oq, os_ = ogemm.per_token_group_quant_fp8(q_contiguous.float().cpu().numpy())
    assert np.array_equal(q8.view(torch.uint8).cpu().numpy(), oq)
    assert np.array_equal(scales.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    # (3) both outputs at once
    c = qkv.clone()
    q_bf16 = torch.empty(m, q_dim, dtype=torch.bfloat16, device=DEV)
    q8c, sc = norm.forward_block_quant(c, q_dim, output_q_a=q_bf16)
    assert torch.equal(q_bf16, q_contiguous) and torch.equal(q8c.view(torch.uint8), q8.view(torch.uint8))


def test_fused_norm_rejects_non_adjacent_inputs():
    from sglang_fluentllm_b200 import layernorm as ln
    norm = ln.FusedRMSNorm(ln.RMSNorm(1536, 1e-6, DEV), ln.RMSNorm(512, 1e-6, DEV))
    q = torch.zeros(4, 1536, dtype=torch.bfloat16, device=DEV)
    kv = torch.zeros(4, 512, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):
        norm(q, kv)
    with pytest.raises(RuntimeError):
        norm.forward_block_quant(torch.zeros(4, 2112, dtype=torch.bfloat16), 1536)   # CPU tensor: no fallback
