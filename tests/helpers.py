"""Test-side input preparation (TEST INFRASTRUCTURE ONLY): FP8 quantisation of random activations / weights the way
the reference's own GEMM tests prepare theirs (3rdparty/deep_gemm/tests/test_core.py:9-25 -- 1x128 groups along k for the
left operand, 128x128 blocks for the right one, scale = amax / 448 with amax clamped at 1e-4), and the reference's
similarity criterion (deep_gemm/utils.py:140-144).  tests/golden/block_fp8.npz pins these against the reference's
functions (tests/test_oracle_golden.py)."""
import torch

FP8_MAX = 448.0


def quantize_lhs_1x128(x: torch.Tensor):
    """x [m, k] -> (e4m3 [m, k], scales f32 [m, k/128])."""
    m, k = x.shape
    assert k % 128 == 0
    groups = x.reshape(m, k // 128, 128).float()
    amax = groups.abs().amax(dim=-1).clamp_min(1e-4)
    q = (groups * (FP8_MAX / amax)[..., None]).to(torch.float8_e4m3fn)
    return q.reshape(m, k), amax / FP8_MAX


def quantize_rhs_128x128(w: torch.Tensor):
    """w [n, k] -> (e4m3 [n, k], scales f32 [ceil(n/128), ceil(k/128)]); partial edge blocks are zero-padded."""
    n, k = w.shape
    nb, kb = -(-n // 128), -(-k // 128)
    padded = w.new_zeros((nb * 128, kb * 128))
    padded[:n, :k] = w
    blocks = padded.reshape(nb, 128, kb, 128).float()
    amax = blocks.abs().amax(dim=(1, 3)).clamp_min(1e-4)
    q = (blocks * (FP8_MAX / amax)[:, None, :, None]).to(torch.float8_e4m3fn)
    return q.reshape(nb * 128, kb * 128)[:n, :k].contiguous(), amax / FP8_MAX


def similarity_diff(x: torch.Tensor, y: torch.Tensor) -> float:
    """1 - 2<x, y> / (|x|^2 + |y|^2): the reference tests' pass criterion (< 1e-3)."""
    x, y = x.double(), y.double()
    return float(1.0 - 2.0 * (x * y).sum() / ((x * x).sum() + (y * y).sum()))
