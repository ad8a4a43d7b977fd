"""GPU: the strided 1x128 activation quantiser (b200_act_quant_1x128_strided) reads a producer's output in place and
gives the same bits as the contiguous quantiser (G-2, bit-exact vs oracle in test_gpu_parity.py) on a .contiguous()
copy -- column slices of a wider row and per-head slices of wider heads."""
import numpy as np
import pytest
import torch

from oracle import gemm as ogemm

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fk():
    from sglang_fluentllm_b200 import _lib, fp8_kernel
    _lib.lib()
    return fp8_kernel


def _same(fk, view):
    xq, xs = fk.per_token_group_quant_fp8_strided(view)
    flat = view.reshape(view.shape[0], -1).contiguous()
    rq, rs = fk.sglang_per_token_group_quant_fp8(flat, 128, column_major_scales=True, scale_tma_aligned=True)
    torch.cuda.synchronize()
    assert xq.shape == rq.shape and xs.shape == rs.shape and xs.stride() == rs.stride()
    assert torch.equal(xq.view(torch.uint8), rq.view(torch.uint8))
    assert torch.equal(xs.view(torch.int32), rs.view(torch.int32))
    oq, osc = ogemm.per_token_group_quant_fp8(flat.float().cpu().numpy())
    assert np.array_equal(xq.view(torch.uint8).cpu().numpy(), oq)
    assert np.array_equal(np.ascontiguousarray(xs.cpu().numpy()).view(np.uint32), osc.view(np.uint32))


@pytest.mark.parametrize("m", [1, 5, 128, 257])
def test_column_slice_of_a_wider_row(fk, m):
    torch.manual_seed(m)
    y = (torch.randn(m, 2112, device=DEV) * 3).to(torch.bfloat16)   # fused qkv_a projection output
    y[0, 128:256] = 0                                                # an all-zero group: scale = eps / 448
    _same(fk, y[:, :1536])
    _same(fk, y[:, 1536:2048])


@pytest.mark.parametrize("m,heads", [(1, 1), (3, 16), (128, 128)])
def test_per_head_slice(fk, m, heads):
    torch.manual_seed(heads)
    o = (torch.randn(m, 1, heads, 512, device=DEV) * 3).to(torch.bfloat16)   # attention output [b, s_q, h, 512]
    _same(fk, o[:, 0, :, :128])
    _same(fk, o[:, 0, :, 384:])


def test_rejects_unaligned_and_cpu(fk):
    y = torch.zeros(4, 2112, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(RuntimeError):
        fk.per_token_group_quant_fp8_strided(y[:, 4:132])            # 8-byte aligned start: not a 16-byte multiple
    with pytest.raises(RuntimeError):
        fk.per_token_group_quant_fp8_strided(torch.zeros(4, 256, dtype=torch.bfloat16))
