"""GPU parity tests of the grouped block-FP8 GEMMs (deep_gemm.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous
/ _masked / _offset, 3rdparty/deep_gemm/deep_gemm/jit_kernels/m_grouped_gemm.py:55-310), through the C ABI.  Checker: the
block-FP8 oracle per group on the same FP8 operands (norm-relative < 3e-3) and the reference test's own criterion,
calc_diff < 1e-3 against the un-quantised product (3rdparty/deep_gemm/tests/test_core.py:87-130)."""
import numpy as np
import pytest
import torch

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from helpers import quantize_lhs_1x128, quantize_rhs_128x128, similarity_diff

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def dg():
    from sglang_fluentllm_b200 import _lib, deep_gemm
    _lib.lib()
    return deep_gemm


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def construct_grouped(dg, num_groups, m, k, n, seed):
    """test_core.py:45-68: per-group bf16 operands, 1x128 / 128x128 FP8 casts."""
    torch.manual_seed(seed)
    x = torch.randn(num_groups, m, k, dtype=torch.bfloat16, device=DEV)
    y = torch.randn(num_groups, n, k, dtype=torch.bfloat16, device=DEV)
    x8 = torch.empty_like(x, dtype=torch.float8_e4m3fn)
    xs = torch.empty(num_groups, m, k // 128, dtype=torch.float32, device=DEV)
    y8 = torch.empty_like(y, dtype=torch.float8_e4m3fn)
    ys = torch.empty(num_groups, (n + 127) // 128, k // 128, dtype=torch.float32, device=DEV)
    for g in range(num_groups):
        x8[g], xs[g] = quantize_lhs_1x128(x[g])
        y8[g], ys[g] = quantize_rhs_128x128(y[g])
    return x, y, x8, xs, y8, ys


def oracle_group(x8, xs, y8, ys):
    return ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(x8.view(torch.uint8).cpu().numpy()), xs.cpu().numpy(),
                                       ofp8.e4m3_to_f32(y8.view(torch.uint8).cpu().numpy()), ys.cpu().numpy())


@pytest.mark.parametrize("num_groups,m,k,n", [(4, 256, 512, 256), (3, 128, 2048, 1792), (8, 128, 7168, 4096)])
def test_grouped_contiguous_vs_oracle(dg, num_groups, m, k, n):
    x, y, x8, xs, y8, ys = construct_grouped(dg, num_groups, m, k, n, seed=num_groups + n)
    # groups laid out back to back in a shuffled order, one block of -1 (skipped) rows in the middle
    order = torch.randperm(num_groups).tolist()
    m_indices = torch.cat([torch.full((m,), g, dtype=torch.int32) for g in order[:1]] +
                          [torch.full((128,), -1, dtype=torch.int32)] +
                          [torch.full((m,), g, dtype=torch.int32) for g in order[1:]]).to(DEV)
    pad8 = torch.zeros(128, k, dtype=torch.float8_e4m3fn, device=DEV)
    pads = torch.ones(128, k // 128, dtype=torch.float32, device=DEV)
    lhs = torch.cat([x8[order[0]], pad8] + [x8[g] for g in order[1:]]).contiguous()
    lhs_s = torch.cat([xs[order[0]], pads] + [xs[g] for g in order[1:]]).contiguous()
    out = torch.full((lhs.shape[0], n), 7.0, dtype=torch.bfloat16, device=DEV)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous((lhs, lhs_s), (y8, ys), out, m_indices, True)
    torch.cuda.synchronize()
    assert torch.all(out[m:m + 128] == 7.0), "a skipped block was written"
    row = 0
    for i, g in enumerate(order):
        if i == 1:
            row += 128
        got = out[row:row + m].float().cpu()
        assert rel(got.numpy(), oracle_group(x8[g], xs[g], y8[g], ys[g])) < 3e-3
        assert similarity_diff(got, (x[g].float() @ y[g].float().t()).cpu()) < 1e-3
        row += m


@pytest.mark.parametrize("num_groups,m,k,n", [(1, 256, 512, 256), (2, 512, 2048, 1792), (4, 256, 7168, 4096)])
def test_grouped_masked_vs_oracle(dg, num_groups, m, k, n):
    x, y, x8, xs, y8, ys = construct_grouped(dg, num_groups, m, k, n, seed=17 * num_groups + n)
    rng = np.random.default_rng(num_groups)
    masked = [int(v) for v in rng.choice([0, 1, 64, 100, 128, 192, m], size=num_groups)]
    masked_m = torch.tensor(masked, dtype=torch.int32, device=DEV)
    out = torch.full((num_groups, m, n), 7.0, dtype=torch.bfloat16, device=DEV)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_masked((x8, xs), (y8, ys), out, masked_m, max(1, int(np.mean(masked))), True)
    torch.cuda.synchronize()
    for g in range(num_groups):
        mm = masked[g]
        assert torch.all(out[g, mm:] == 7.0), "rows past masked_m were written"
        if mm == 0:
            continue
        got = out[g, :mm].float().cpu()
        assert rel(got.numpy(), oracle_group(x8[g, :mm], xs[g, :mm], y8[g], ys[g])) < 3e-3
        assert similarity_diff(got, (x[g, :mm].float() @ y[g].float().t()).cpu()) < 1e-3


@pytest.mark.parametrize("lens,k,n", [([5], 512, 256), ([130, 0, 1, 127, 64], 512, 256), ([3, 250, 129, 2, 77, 300], 2048, 1792),
                                      ([100, 37, 0, 0, 260, 1, 4, 7, 127, 126], 7168, 4096)])
def test_grouped_offset_vs_oracle(dg, lens, k, n):
    """Offset layout: groups of arbitrary (also zero, also non-multiple-of-4) row counts back to back, plus a tail of rows
    no group owns."""
    num_groups, m_max = len(lens), max(max(lens), 1)
    x, y, x8, xs, y8, ys = construct_grouped(dg, num_groups, m_max, k, n, seed=sum(lens) + n)
    tail = 9
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    m_sum = int(off[-1]) + tail
    pad8 = torch.zeros(tail, k, dtype=torch.float8_e4m3fn, device=DEV)
    pads = torch.ones(tail, k // 128, dtype=torch.float32, device=DEV)
    lhs = torch.cat([x8[g, :lens[g]] for g in range(num_groups)] + [pad8]).contiguous()
    lhs_s = torch.cat([xs[g, :lens[g]] for g in range(num_groups)] + [pads]).contiguous()
    out = torch.full((m_sum, n), 7.0, dtype=torch.bfloat16, device=DEV)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_offset((lhs, lhs_s), (y8, ys), out, torch.from_numpy(off).to(DEV), True)
    torch.cuda.synchronize()
    assert torch.all(out[int(off[-1]):] == 7.0), "rows past the last group were written"
    for g in range(num_groups):
        if lens[g] == 0:
            continue
        got = out[int(off[g]):int(off[g + 1])].float().cpu()
        assert rel(got.numpy(), oracle_group(x8[g, :lens[g]], xs[g, :lens[g]], y8[g], ys[g])) < 3e-3
        if lens[g] >= 16:
            assert similarity_diff(got, (x[g, :lens[g]].float() @ y[g].float().t()).cpu()) < 1e-3


def test_grouped_layouts_agree(dg):
    """Size-independent property: the three layouts are the same arithmetic -- identical bits for the same rows."""
    num_groups, m, k, n = 3, 128, 1024, 512
    x, y, x8, xs, y8, ys = construct_grouped(dg, num_groups, m, k, n, seed=77)
    lhs, lhs_s = x8.reshape(num_groups * m, k), xs.reshape(num_groups * m, k // 128)
    o1 = torch.zeros(num_groups * m, n, dtype=torch.bfloat16, device=DEV)
    o2 = torch.zeros(num_groups, m, n, dtype=torch.bfloat16, device=DEV)
    o3 = torch.zeros_like(o1)
    idx = torch.arange(num_groups, dtype=torch.int32, device=DEV).repeat_interleave(m)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous((lhs, lhs_s), (y8, ys), o1, idx)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_masked((x8, xs), (y8, ys), o2, torch.full((num_groups,), m, dtype=torch.int32, device=DEV), m)
    dg.m_grouped_gemm_fp8_fp8_bf16_nt_offset((lhs, lhs_s), (y8, ys), o3,
                                             torch.arange(num_groups + 1, dtype=torch.int32, device=DEV) * m)
    plain = torch.zeros_like(o1)
    for g in range(num_groups):
        dg.gemm_fp8_fp8_bf16_nt((x8[g], dg.get_col_major_tma_aligned_tensor(xs[g])), (y8[g], ys[g]), plain[g * m:(g + 1) * m])
    torch.cuda.synchronize()
    assert torch.equal(o1.view(torch.int16), o2.view(num_groups * m, n).view(torch.int16))
    assert torch.equal(o1.view(torch.int16), o3.view(torch.int16))
    # the dense entry may split K (other summation order): same numbers up to fp32 reassociation and one bf16 rounding
    assert rel(o1.float().cpu().numpy(), plain.float().cpu().numpy()) < 2e-3
