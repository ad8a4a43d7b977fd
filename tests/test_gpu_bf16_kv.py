"""GPU parity tests of the bf16-KV MLA decode kernel (flash_mla_with_kvcache of both reference
forks), through the C ABI.  Checker: oracle/mla.py::flash_mla_with_kvcache (float64 attention over the gathered
pages) on the same seeded inputs, with the reference's own criterion cos_diff < 1e-5 for out and lse
(3rdparty/flashmla-fp8/tests/test_flash_mla.py:31-37,104-105) plus a norm-relative bound (bf16 rounding of P and of
the output, ~2e-3), and -- at the BASELINE shape -- page-permutation invariance bit for bit."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mla as omla

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def fm():
    from sglang_fluentllm_b200 import _lib, flash_mla_fp8
    _lib.lib()
    return flash_mla_fp8


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def cos_diff(x, y):
    x, y = x.astype(np.float64), y.astype(np.float64)
    return 1 - 2 * (x * y).sum() / max((x * x + y * y).sum(), 1e-12)


def test_umma_bf16_mn_major_two_chunk_descriptor():
    """P.V operand forms of the bf16 kernel: A = V^T MN-major SW128 spanning two 64-dim chunks (LBO = chunk stride),
    B = thread-written bf16 P[q][tok] K-major SW128 in two 64-token chunks."""
    from sglang_fluentllm_b200 import _lib
    L = _lib.lab()  # the probe library (include/b200_lab.h)
    for n in (16, 32, 64):
        err = ctypes.c_float(-1)
        rc = L.b200_umma_selftest(5, n, ctypes.byref(err))
        assert rc == 0, L.b200_lab_last_error()
        assert err.value < 1e-5, "variant 5 n=%d rel_err=%g" % (n, err.value)


def make_case(b, h, s_q, seqlens, seed=0, shuffle=True, nan_pad=True):
    torch.manual_seed(seed)
    d = 576
    seqlens = torch.as_tensor(seqlens, dtype=torch.int32)
    max_blocks = int((int(seqlens.max()) + 63) // 64) + 1
    n_pages = b * max_blocks + 1
    q = torch.randn(b, s_q, h, d, dtype=torch.bfloat16, device=DEV)
    kv = torch.randn(n_pages, 64, 1, d, dtype=torch.bfloat16, device=DEV)
    perm = torch.randperm(n_pages - 1) + 1 if shuffle else torch.arange(1, n_pages)
    bt = perm[: b * max_blocks].reshape(b, max_blocks).to(torch.int32)
    if nan_pad:  # the reference test poisons everything past seqlen (test_flash_mla.py:62-65)
        for i in range(b):
            sl = int(seqlens[i])
            blk, off = sl // 64, sl % 64
            for j in range(blk, max_blocks):
                kv[int(bt[i, j]), (off if j == blk else 0):] = float("nan")
    return q, kv, bt, seqlens


def run(fm, q, kv, bt, sl, causal):
    b, s_q, h, _ = q.shape
    sld = sl.to(DEV)
    meta, ns = fm.get_mla_metadata(sld, s_q * h, 1)
    out, lse = fm.flash_mla_with_kvcache(q, kv, bt.to(DEV), sld, 512, meta, ns, causal=causal)
    torch.cuda.synchronize()
    return out.float().cpu().numpy(), lse.cpu().numpy(), ns.cpu().numpy()


CASES = [
    ("one-page", 1, 16, 1, [17], False),
    ("exact-page", 2, 16, 1, [64, 128], False),
    ("ragged-h16", 4, 16, 1, [300, 64, 1, 517], False),
    ("ragged-h32", 3, 32, 1, [129, 700, 65], False),
    ("ragged-h64", 3, 64, 1, [129, 700, 65], False),        # two q blocks per request
    ("ragged-h128", 2, 128, 1, [333, 1025], False),         # four q blocks
    ("m48", 2, 48, 1, [500, 130], False),                   # second q block half valid
    ("sq2-causal-h16", 3, 16, 2, [300, 130, 64], True),
    ("sq4-causal-h16", 2, 16, 4, [300, 65], True),
    ("sq2-causal-h64", 2, 64, 2, [257, 64], True),
    ("sq2-noncausal", 2, 16, 2, [200, 90], False),
]


@pytest.mark.parametrize("name,b,h,s_q,seqlens,causal", CASES, ids=[c[0] for c in CASES])
def test_bf16_decode_vs_oracle(fm, name, b, h, s_q, seqlens, causal):
    q, kv, bt, sl = make_case(b, h, s_q, seqlens)
    out, lse, _ = run(fm, q, kv, bt, sl, causal)
    assert not np.isnan(out).any(), "NaN leaked from the poisoned padding rows"
    ro, rl = omla.flash_mla_with_kvcache(q.float().cpu().numpy(), kv.float().cpu().numpy(), bt.numpy(), sl.numpy(),
                                         causal=causal)
    assert cos_diff(out, ro) < 1e-5      # the reference's own criterion
    assert cos_diff(lse, rl) < 1e-5
    assert rel(out, ro) < 4e-3           # bf16 rounding of P and of the output
    assert np.allclose(lse, rl, rtol=1e-4, atol=1e-3)
    row_err = np.linalg.norm(out - ro, axis=-1) / np.maximum(np.linalg.norm(ro, axis=-1), 1e-20)
    assert row_err.max() < 2e-2


def test_bf16_decode_split_long_sequences(fm):
    """Long ragged sequences: requests are cut over many scheduler parts (fp32 partials + combine kernel)."""
    q, kv, bt, sl = make_case(3, 16, 1, [20000, 9000, 130])
    out, lse, ns = run(fm, q, kv, bt, sl, False)
    assert (np.diff(ns) > 1).any(), "expected at least one split request"
    ro, rl = omla.flash_mla_with_kvcache(q.float().cpu().numpy(), kv.float().cpu().numpy(), bt.numpy(), sl.numpy())
    assert cos_diff(out, ro) < 1e-5 and rel(out, ro) < 4e-3
    assert np.allclose(lse, rl, rtol=1e-4, atol=1e-3)


def test_bf16_reference_moves_under_growing_logits(fm):
    """Keys whose norm grows along the sequence push the logits up tile after tile: the lazy softmax reference has to
    move (O rescaled in TMEM) several times; the result must stay exact and bit-reproducible."""
    b, h, seq = 2, 16, 1536
    q, kv, bt, sl = make_case(b, h, 1, [seq, seq - 200], nan_pad=False, seed=5)
    scale = torch.ones(kv.shape[0], 64, device=DEV)
    for i in range(b):
        t = torch.arange(seq, device=DEV)
        pages = bt[i].to(DEV).long()[t // 64]
        scale[pages, t % 64] = torch.exp2(t.float() / 96.0)
    kv = (kv.float() * scale.view(-1, 64, 1, 1) * 0.05).to(torch.bfloat16)
    out, lse, _ = run(fm, q, kv, bt, sl, False)
    ro, rl = omla.flash_mla_with_kvcache(q.float().cpu().numpy(), kv.float().cpu().numpy(), bt.numpy(), sl.numpy())
    assert np.isfinite(out).all() and np.isfinite(lse).all()
    assert cos_diff(out, ro) < 1e-5 and rel(out, ro) < 4e-3
    assert np.allclose(lse, rl, rtol=1e-4, atol=2e-3)
    out2, lse2, _ = run(fm, q, kv, bt, sl, False)
    assert np.array_equal(out, out2) and np.array_equal(lse, lse2)


def test_bf16_full_size_properties(fm):
    """BASELINE shape (b = 128, seq = 4096, h = 16): page-permutation invariance bit for bit, and out / lse against
    torch attention for a sample of requests."""
    b, h, seq = 128, 16, 4096
    q, kv, bt, sl = make_case(b, h, 1, [seq] * b, nan_pad=False)
    out1, lse1, _ = run(fm, q, kv, bt, sl, False)
    n_pages = kv.shape[0]
    perm = torch.randperm(n_pages - 1) + 1
    inv = torch.zeros(n_pages, dtype=torch.int64)
    inv[perm] = torch.arange(1, n_pages)
    src = torch.cat([torch.zeros(1, dtype=torch.int64), inv[1:]]).to(DEV)  # new page p holds old page inv[p]
    kv2 = kv[src].contiguous()
    bt2 = perm[bt.long() - 1].to(torch.int32)
    out2, lse2, _ = run(fm, q, kv2, bt2, sl, False)
    assert np.array_equal(out1, out2) and np.array_equal(lse1, lse2)
    for i in (0, 41, 127):
        pages = bt[i, : seq // 64].long().to(DEV)
        k = kv[pages].view(-1, 576).float()
        logits = (q[i, 0].float() @ k.t()) * 576 ** -0.5
        assert np.allclose(lse1[i, :, 0], torch.logsumexp(logits, dim=-1).cpu().numpy(), rtol=1e-4, atol=1e-3)
        assert rel(out1[i, 0], (torch.softmax(logits, -1) @ k[:, :512]).cpu().numpy()) < 4e-3
