"""GPU parity tests of the grouped block-FP8 GEMMs (deep_gemm.m_grouped_gemm_fp8_fp8_bf16_nt_contiguous
/ _masked, 3rdparty/deep_gemm/deep_gemm/jit_kernels/m_grouped_gemm.py:55-215), through the C ABI.  Checker: the
block-FP8 oracle per group on the same FP8 operands (norm-relative < 3e-3) and the reference test's own criterion,
calc_diff < 1e-3 against the un-quantised product (3rdparty/deep_gemm/tests/test_core.py:87-130)."""
import numpy as np
import pytest
import torch

from oracle import fp8 as ofp8
from oracle import gemm as ogemm

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
        x8[g], xs[g] = dg.per_token_cast_to_fp8(x[g])
        y8[g], ys[g] = dg.per_block_cast_to_fp8(y[g])
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
        assert dg.calc_diff(got, (x[g].float() @ y[g].float().t()).cpu()) < 1e-3
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
        assert dg.calc_diff(got, (x[g, :mm].float() @ y[g].float().t()).cpu()) < 1e-3
