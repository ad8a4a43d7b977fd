"""GPU parity tests (run with -m gpu on the B200 box).  Every call goes through the C ABI (ctypes -> libb200decode.so);
the checker is the CPU oracle under oracle/ (pinned by tests/test_oracle_golden.py) on the same seeded inputs.

Tolerances
  * integer / byte work (scheduler metadata, e4m3 codes, fp32 scale bits, bf16 rope bits, block-table gather):
    bit-exact.
  * block-FP8 GEMM: the reference's criterion, calc_diff < 1e-3 (3rdparty/deep_gemm/tests/test_core.py:71-72), and
    norm-relative < 3e-3 against the fp32 oracle (bf16 output rounding is ~1.7e-3).
  * MLA decode: north_star's 1e-2 relative against the reference-faithful page-wise FP8-P recurrence (oracle/mla.py
    ::flash_mla_ckv_fp8_per_token, splitkv_mla.cu:377-458) -- measured ~1e-3: the kernels produce the reference's own
    per-page e4m3 codes of P and carry the page weight as an e4m3 hi + lo pair; lse to 1e-3; and the reference's own
    criterion rel < 5e-2 against fp32 attention (tests/test_flash_mla_ckv_fp8_per_token.py:32-39,119-120; ~2e-2 is the
    e4m3 noise of P on zero-mean V, common to the reference and to us).  See DESIGN.md "Numerics".
"""
import ctypes

import os

import numpy as np
import pytest
import torch

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from oracle import mla as omla
from oracle import pool as opool
from helpers import quantize_lhs_1x128, quantize_rhs_128x128, similarity_diff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from sglang_fluentllm_b200 import _lib, deep_gemm, flash_mla_fp8, fp8_kernel
    _lib.lib()
    return _lib, flash_mla_fp8, deep_gemm, fp8_kernel


DEV = "cuda:0"


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


# ------------------------------------------------------------------------------------------------------------------
def test_umma_descriptor_selftest(mods):
    """The hand-built UMMA smem/instruction descriptors of every operand layout the kernels use."""
    L = mods[0].lab()  # the probe library (include/b200_lab.h), not the product
    for variant in (0, 1, 3):  # K-major fp8, K-major bf16, MN-major V^T x thread-written P
        for n in (16, 64):
            err = ctypes.c_float(-1)
            rc = L.b200_umma_selftest(variant, n, ctypes.byref(err))
            assert rc == 0, L.b200_lab_last_error()
            assert err.value < 1e-5, "variant %d n=%d rel_err=%g" % (variant, n, err.value)


@pytest.mark.parametrize("batch,hpk", [(128, 16), (128, 128), (37, 64), (1, 16), (300, 32)])
def test_metadata_bit_exact(mods, batch, hpk):
    _lib, fm, _, _ = mods
    rng = np.random.default_rng(batch)
    seqlens = rng.integers(1, 9000, size=batch).astype(np.int32) if batch not in (128,) else np.full(batch, 4096, np.int32)
    meta, ns = fm.get_mla_metadata(torch.from_numpy(seqlens).to(DEV), hpk, 1)
    om, ons = omla.get_mla_metadata(seqlens, hpk, 1, _lib.lib().b200_sm_count())
    assert np.array_equal(meta.cpu().numpy()[:, :5], om[:, :5])
    assert np.array_equal(ns.cpu().numpy(), ons)
    # the scheduler with the per-request cost the B200 backend passes (an extension; same algorithm, same format)
    for ovh in (fm.tuned_overhead_blocks(hpk), 3, 40):
        meta, ns = fm.get_mla_metadata(torch.from_numpy(seqlens).to(DEV), hpk, 1, fixed_overhead_blocks=ovh)
        om, ons = omla.get_mla_metadata(seqlens, hpk, 1, _lib.lib().b200_sm_count(), fixed_overhead_blocks=ovh)
        assert np.array_equal(meta.cpu().numpy()[:, :5], om[:, :5]), ovh
        assert np.array_equal(ns.cpu().numpy(), ons), ovh
    # ... and with the number of parts the backend asks for (33-64 q rows run on CTA pairs: half as many parts): the
    # oracle gets there through an SM count that gives the same number under the reference's rule
    parts = fm.tuned_num_sm_parts(hpk)
    if parts is not None:
        ovh = fm.tuned_overhead_blocks(hpk)
        meta, ns = fm.get_mla_metadata(torch.from_numpy(seqlens).to(DEV), hpk, 1, fixed_overhead_blocks=ovh, num_sm_parts=parts)
        assert meta.shape[0] == parts
        om, ons = omla.get_mla_metadata(seqlens, hpk, 1, parts * ((hpk + 63) // 64), fixed_overhead_blocks=ovh)
        assert np.array_equal(meta.cpu().numpy()[:, :5], om[:, :5]) and np.array_equal(ns.cpu().numpy(), ons)


def test_quantize_q_bit_exact(mods):
    fm = mods[1]
    torch.manual_seed(0)
    q = torch.randn(5, 2, 16, 576, dtype=torch.bfloat16, device=DEV) * torch.rand(5, 2, 16, 1, device=DEV).to(torch.bfloat16) * 8
    q[0, 0, 3] = 0
    nope, scale, rope = fm.quantize_ckv_per_token_head(q, 512)
    oc, osc, orp = omla.quantize_ckv_per_token_head(q.float().cpu().numpy(), 512)
    assert nope.shape == (5, 2, 16, 512) and scale.shape == (5, 2, 16, 1) and rope.shape == (5, 2, 16, 64)
    assert np.array_equal(nope.cpu().numpy(), oc)
    assert np.array_equal(scale.cpu().numpy().view(np.uint32), osc.view(np.uint32))
    assert np.array_equal(rope.float().cpu().numpy(), orp)


def test_quantize_and_cache_k_and_dequant_bit_exact(mods):
    fm = mods[1]
    torch.manual_seed(1)
    slots = 64 * 9
    key = (torch.randn(77, 1, 576, device=DEV) * 3).to(torch.bfloat16)
    loc = (torch.randperm(slots - 64)[:77] + 64).to(torch.int32).to(DEV)
    kl = torch.zeros(slots, 1, 512, dtype=torch.uint8, device=DEV)
    ks = torch.zeros(slots, 1, 1, dtype=torch.float32, device=DEV)
    kr = torch.zeros(slots, 1, 64, dtype=torch.bfloat16, device=DEV)
    fm.quantize_and_cache_k(key, kl, ks, kr, loc, 512)
    okl, oks, okr = np.zeros((slots, 1, 512), np.uint8), np.zeros((slots, 1, 1), np.float32), np.zeros((slots, 1, 64), np.float32)
    omla.quantize_and_cache_k(key.float().cpu().numpy(), okl, oks, okr, loc.cpu().numpy())
    assert np.array_equal(kl.cpu().numpy(), okl)
    assert np.array_equal(ks.cpu().numpy().view(np.uint32), oks.view(np.uint32))
    assert np.array_equal(kr.float().cpu().numpy(), okr)
    # indexed de-quantisation, incl. repeated indices, empty request and the CUDA-graph max_num_tokens form
    idx = torch.cat([loc[:40], loc[:3]]).contiguous()
    for mx in (-1, 256):
        dl, dr = fm.dequantize_ckv_fused_indexed(kl, kr, ks, idx, mx)
        ol, orr = omla.dequantize_ckv_fused_indexed(okl, okr, oks, idx.cpu().numpy())
        assert np.array_equal(dl.float().cpu().numpy(), ol) and np.array_equal(dr.float().cpu().numpy(), orr)
    dl, dr = fm.dequantize_ckv_fused_indexed(kl, kr, ks, idx[:0].contiguous())
    assert dl.shape == (0, 1, 512) and dr.shape == (0, 1, 64)


def test_quantize_and_cache_k_strided_rows(mods):
    """key rows as a strided view (the kv columns of the fused qkv_a projection output): same bytes as a contiguous copy."""
    fm = mods[1]
    torch.manual_seed(2)
    wide = (torch.randn(33, 2112, device=DEV) * 2).to(torch.bfloat16)
    key = wide[:, 1536:].unsqueeze(1)
    assert not key.is_contiguous()
    loc = (torch.randperm(64 * 3)[:33] + 64).to(torch.int32).to(DEV)
    outs = []
    for k in (key, key.contiguous()):
        kl = torch.zeros(64 * 4, 1, 512, dtype=torch.uint8, device=DEV)
        ks = torch.zeros(64 * 4, 1, 1, dtype=torch.float32, device=DEV)
        kr = torch.zeros(64 * 4, 1, 64, dtype=torch.bfloat16, device=DEV)
        fm.quantize_and_cache_k(k, kl, ks, kr, loc, 512)
        outs.append((kl, ks, kr))
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_quantiser_matches_reference_fallback_golden(mods, golden_dir):
    """CUDA quantiser vs the reference's own torch fall-back output (tests/golden/pool_per_token_head.npz)."""
    import os
    fm = mods[1]
    g = np.load(os.path.join(golden_dir, "pool_per_token_head.npz"))
    key = torch.from_numpy(g["key"]).to(torch.bfloat16).to(DEV)
    loc = torch.from_numpy(g["loc"]).to(DEV)
    slots = g["k_lora"].shape[0]
    kl = torch.zeros(slots, 1, 512, dtype=torch.uint8, device=DEV)
    ks = torch.zeros(slots, 1, 1, dtype=torch.float32, device=DEV)
    kr = torch.zeros(slots, 1, 64, dtype=torch.bfloat16, device=DEV)
    fm.quantize_and_cache_k(key, kl, ks, kr, loc, 512)
    nz = g["k_scale"][g["loc"], 0, 0] > 1e-20
    got_s = ks.cpu().numpy()[g["loc"], 0, 0]
    assert np.allclose(got_s[nz], g["k_scale"][g["loc"], 0, 0][nz], rtol=1e-6)
    mism = (kl.cpu().numpy()[g["loc"]] != g["k_lora"][g["loc"]])[nz].mean()
    assert mism < 5e-3
    dl, dr = fm.dequantize_ckv_fused_indexed(torch.from_numpy(g["k_lora"]).to(DEV), torch.from_numpy(g["k_rope"]).to(torch.bfloat16).to(DEV),
                                             torch.from_numpy(g["k_scale"]).to(DEV), loc)
    assert np.array_equal(dl.float().cpu().numpy(), g["deq_lora"]) and np.array_equal(dr.float().cpu().numpy(), g["deq_rope"])


@pytest.mark.parametrize("m,k", [(1, 128), (33, 512), (128, 7168), (257, 2048)])
def test_act_quant_bit_exact(mods, m, k):
    fk = mods[3]
    torch.manual_seed(m)
    x = (torch.randn(m, k, device=DEV) * 2).to(torch.bfloat16)
    if m > 2:
        x[2, :128] = 0
    xq, xs = fk.sglang_per_token_group_quant_fp8(x, 128, column_major_scales=True, scale_tma_aligned=True)
    oq, os_ = ogemm.per_token_group_quant_fp8(x.float().cpu().numpy())
    assert xs.stride(0) == 1 and xs.shape == (m, k // 128)
    assert np.array_equal(xq.view(torch.uint8).cpu().numpy(), oq)
    assert np.array_equal(xs.cpu().numpy().view(np.uint32), os_.view(np.uint32))


def test_act_quant_golden(mods, golden_dir):
    import os
    fk = mods[3]
    g = np.load(os.path.join(golden_dir, "block_fp8.npz"))
    x = torch.from_numpy(g["x"]).to(torch.bfloat16).to(DEV)
    xq, xs = fk.sglang_per_token_group_quant_fp8(x, 128, column_major_scales=True, scale_tma_aligned=True)
    assert np.array_equal(xq.view(torch.uint8).cpu().numpy(), g["xq"])
    zero_group = np.abs(g["x"]).reshape(g["xs"].shape[0], g["xs"].shape[1], 128).max(-1) < 1e-10
    assert np.array_equal(xs.cpu().numpy()[~zero_group].view(np.uint32), g["xs"][~zero_group].view(np.uint32))


GEMM_SHAPES = [(16, 128, 128), (1, 64, 128), (32, 256, 512), (100, 192, 256), (300, 256, 384), (128, 2112, 7168),
               (128, 7168, 2048), (64, 4096, 7168), (7, 2112, 7168), (128, 24576, 1536),
               # n a multiple of 192: the 192-row weight tile (tiles that start in the middle of a 128-row scale block),
               # with and without split-K (`B200_GEMM_BN=192 pytest -k gemm` forces it on every shape it fits)
               (128, 3072, 1536), (200, 768, 2048), (130, 1920, 7168), (128, 384, 16384)]


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
def test_fp8_gemm_vs_oracle(mods, m, n, k):
    dg = mods[2]
    torch.manual_seed(m * 7 + n)
    x = torch.randn(m, k, dtype=torch.bfloat16, device=DEV)
    w = torch.randn(n, k, dtype=torch.bfloat16, device=DEV)
    a8, a_s = quantize_lhs_1x128(x)
    w8, w_s = quantize_rhs_128x128(w)
    c = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
    for _ in range(2):  # second call checks the self-cleaning split-K workspace
        dg.gemm_fp8_fp8_bf16_nt((a8, dg.get_col_major_tma_aligned_tensor(a_s)), (w8, w_s), c, True)
    torch.cuda.synchronize()
    ref = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(a8.view(torch.uint8).cpu().numpy()), a_s.cpu().numpy(),
                                      ofp8.e4m3_to_f32(w8.view(torch.uint8).cpu().numpy()), w_s.cpu().numpy())
    got = c.float().cpu().numpy()
    assert rel(got, ref) < 3e-3
    assert similarity_diff(c.float().cpu(), torch.from_numpy(ref)) < 1e-3  # test_core.py:71-72
    # and against the un-quantised product, the reference test's own comparison
    assert similarity_diff(c.float().cpu(), (x.float() @ w.float().t()).cpu()) < 1e-3


@pytest.mark.parametrize("m", [256, 4096])
@pytest.mark.parametrize("n,k", [(7168, 2048), (4096, 7168)])
def test_fp8_gemm_baseline_microbench_points(mods, m, n, k):
    """The BASELINE.json GEMM shapes at the decode batch 256 and at the compute-bound m = 4096: calc_diff < 1e-3 against
    the oracle's two-level accumulation, evaluated on sampled output rows (the fp32 oracle is a CPU matmul)."""
    dg = mods[2]
    torch.manual_seed(m + n)
    x = torch.randn(m, k, dtype=torch.bfloat16, device=DEV)
    w = torch.randn(n, k, dtype=torch.bfloat16, device=DEV)
    a8, a_s = quantize_lhs_1x128(x)
    w8, w_s = quantize_rhs_128x128(w)
    c = torch.full((m, n), float("nan"), dtype=torch.bfloat16, device=DEV)
    dg.gemm_fp8_fp8_bf16_nt((a8, dg.get_col_major_tma_aligned_tensor(a_s)), (w8, w_s), c, True)
    torch.cuda.synchronize()
    assert not torch.isnan(c).any()
    rows = torch.arange(0, m, max(1, m // 96))[:96]          # rows spread over every 128-row tile class
    ref = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(a8[rows.to(DEV)].view(torch.uint8).cpu().numpy()),
                                      a_s[rows.to(DEV)].cpu().numpy(),
                                      ofp8.e4m3_to_f32(w8.view(torch.uint8).cpu().numpy()), w_s.cpu().numpy())
    got = c[rows.to(DEV)].float().cpu().numpy()
    assert rel(got, ref) < 3e-3
    assert similarity_diff(torch.from_numpy(got), torch.from_numpy(ref)) < 1e-3
    # every row against the un-quantised product on the GPU (the reference test's own comparison)
    assert similarity_diff(c.float().cpu(), (x.float() @ w.float().t()).cpu()) < 1e-3


def test_fp8_gemm_golden(mods, golden_dir):
    import os
    dg = mods[2]
    g = np.load(os.path.join(golden_dir, "block_fp8.npz"))
    a8 = torch.from_numpy(g["xq"]).to(DEV).view(torch.float8_e4m3fn)
    w8 = torch.from_numpy(g["wq"]).to(DEV).view(torch.float8_e4m3fn)
    c = torch.empty(a8.shape[0], w8.shape[0], dtype=torch.bfloat16, device=DEV)
    dg.gemm_fp8_fp8_bf16_nt((a8, torch.from_numpy(g["xs"]).to(DEV)), (w8, torch.from_numpy(g["ws"]).to(DEV)), c, False)
    assert rel(c.float().cpu().numpy(), g["c"]) < 3e-3


# ------------------------------------------------------------------------------------------------------------------
def make_case(fm, b, h, s_q, seqlens, seed=0, shuffle=True, nan_pad=True):
    torch.manual_seed(seed)
    d, dv = 576, 512
    seqlens = torch.as_tensor(seqlens, dtype=torch.int32)
    max_blocks = int((int(seqlens.max()) + 63) // 64) + 1
    n_pages = b * max_blocks + 1
    q = torch.randn(b, s_q, h, d, dtype=torch.bfloat16, device=DEV)
    kv = torch.randn(n_pages * 64, 1, d, dtype=torch.bfloat16, device=DEV)
    perm = torch.randperm(n_pages - 1) + 1 if shuffle else torch.arange(1, n_pages)
    block_table = perm[: b * max_blocks].reshape(b, max_blocks).to(torch.int32)
    if nan_pad:  # reference test poisons everything past seqlen (test_flash_mla_ckv_fp8_per_token.py:64-67)
        kvv = kv.view(n_pages, 64, d)
        for i in range(b):
            sl = int(seqlens[i])
            blk, off = sl // 64, sl % 64
            for j in range(blk, max_blocks):
                kvv[int(block_table[i, j]), (off if j == blk else 0):] = float("nan")
    kl = torch.zeros(n_pages * 64, 1, dv, dtype=torch.uint8, device=DEV)
    ks = torch.zeros(n_pages * 64, 1, 1, dtype=torch.float32, device=DEV)
    kr = torch.zeros(n_pages * 64, 1, d - dv, dtype=torch.bfloat16, device=DEV)
    fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=DEV), dv)
    qn, qs, qr = fm.quantize_ckv_per_token_head(q, dv)
    return dict(qn=qn, qr=qr, qs=qs, kl=kl.view(n_pages, 64, 1, dv), kr=kr.view(n_pages, 64, 1, d - dv),
                ks=ks.view(n_pages, 64, 1, 1), bt=block_table, sl=seqlens)


def run_mla(fm, c, h, s_q, causal, overhead=None, parts=None):
    sl = c["sl"].to(DEV)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, fixed_overhead_blocks=overhead, num_sm_parts=parts)
    out, lse = fm.flash_mla_ckv_fp8_per_token(c["qn"], c["qr"], c["kl"], c["kr"], c["qs"], c["ks"], c["bt"].to(DEV), sl,
                                              512, meta, ns, causal=causal)
    torch.cuda.synchronize()
    args = (c["qn"].cpu().numpy(), c["qr"].float().cpu().numpy(), c["kl"].cpu().numpy(), c["kr"].float().cpu().numpy(),
            c["qs"].cpu().numpy(), c["ks"].cpu().numpy(), c["bt"].numpy(), c["sl"].numpy())
    return out.float().cpu().numpy(), lse.cpu().numpy(), args, meta.cpu().numpy(), ns.cpu().numpy()


MLA_CASES = [
    ("one-page", 1, 16, 1, [17], False),
    ("exact-page", 2, 16, 1, [64, 128], False),
    ("ragged-h16", 4, 16, 1, [300, 64, 1, 517], False),
    ("ragged-h32", 3, 32, 1, [129, 700, 65], False),
    ("ragged-h64", 3, 64, 1, [129, 700, 65], False),
    ("ragged-h128", 2, 128, 1, [333, 1025], False),
    ("sq2-causal-h16", 3, 16, 2, [300, 130, 64], True),
    ("sq4-causal-h16", 2, 16, 4, [300, 65], True),
    ("sq2-causal-h64", 2, 64, 2, [257, 64], True),
    ("sq2-noncausal", 2, 16, 2, [200, 90], False),
    # M > 64: the CTA-pair kernel (two softmax groups, four S / P tiles, per-request epilogue staging)
    ("pair-ragged-b5", 5, 128, 1, [1, 64, 129, 2000, 700], False),
    ("pair-m96", 2, 96, 1, [500, 130], False),                 # second CTA of the pair has 32 valid q rows
    ("pair-sq2-causal-h128", 2, 128, 2, [300, 65], True),      # M = 256: two pairs per request
    ("pair-sq4-causal-h32", 2, 32, 4, [300, 200], True),       # M = 128 from s_q = 4
    ("pair-sq2-m192", 2, 96, 2, [260, 131], True),             # M = 192: a full pair and a half pair
]


@pytest.mark.parametrize("schedule", ["reference", "tuned"])
@pytest.mark.parametrize("name,b,h,s_q,seqlens,causal", MLA_CASES, ids=[c[0] for c in MLA_CASES])
def test_mla_decode_vs_oracles(mods, name, b, h, s_q, seqlens, causal, schedule):
    """Under the reference's tile schedule (the default of get_mla_metadata) and under the one the B200 backend asks
    for (tuned_overhead_blocks): the faithful oracle follows whichever metadata it is given."""
    fm = mods[1]
    c = make_case(fm, b, h, s_q, seqlens)
    tuned = schedule != "reference"
    out, lse, args, meta, ns = run_mla(fm, c, h, s_q, causal, overhead=fm.tuned_overhead_blocks(h * s_q) if tuned else None,
                                       parts=fm.tuned_num_sm_parts(h * s_q) if tuned else None)
    assert not np.isnan(out).any(), "NaN leaked from the poisoned padding rows"
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args, causal=causal)
    fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta, ns, causal=causal)
    assert np.allclose(lse, sl, rtol=1e-3, atol=1e-3)
    assert np.allclose(lse, fl, rtol=1e-3, atol=1e-3)
    assert rel(out, so) < 5e-2          # the reference's own pass criterion
    assert rel(out, fo) < 1e-2, rel(out, fo)   # north_star tolerance vs the reference-faithful FP8-P recurrence
    frow = np.linalg.norm(out - fo, axis=-1) / np.maximum(np.linalg.norm(fo, axis=-1), 1e-20)
    assert frow.max() < 2e-2, frow.max()       # and no single (request, head) row far off it
    # per-row check: no single (request, head) row may be wrong while the norm hides it
    row_err = np.linalg.norm(out - so, axis=-1) / np.maximum(np.linalg.norm(so, axis=-1), 1e-20)
    assert row_err.max() < 0.2


def test_mla_decode_split_long_sequences(mods):
    """Long ragged sequences force the tile scheduler to split requests over many SM parts (combine path)."""
    fm = mods[1]
    c = make_case(fm, 3, 16, 1, [20000, 9000, 130])
    out, lse, args, meta, ns = run_mla(fm, c, 16, 1, False)
    assert (np.diff(ns) > 1).any(), "expected at least one split request"
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args)
    assert np.allclose(lse, sl, rtol=1e-3, atol=1e-3)
    assert rel(out, so) < 5e-2


def test_mla_pair_split_long_sequences(mods):
    """h = 128: 74 scheduler parts over two long requests -> every part boundary falls inside a request (fp32 partials +
    combine), several tiles per part (steady-state ping-pong of the two softmax groups)."""
    fm = mods[1]
    c = make_case(fm, 2, 128, 1, [9000, 4100])
    out, lse, args, meta, ns = run_mla(fm, c, 128, 1, False)
    assert (np.diff(ns) > 8).all(), "expected both requests to be split over many parts"
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args)
    assert np.allclose(lse, sl, rtol=1e-3, atol=1e-3)
    assert rel(out, so) < 5e-2
    row_err = np.linalg.norm(out - so, axis=-1) / np.maximum(np.linalg.norm(so, axis=-1), 1e-20)
    assert row_err.max() < 0.2


@pytest.mark.parametrize("b,h,s_q,seqlens,causal", [
    (2, 128, 1, [9000, 4100], False),        # every part boundary inside a request: up to ~50 parts per request
    (40, 128, 1, [700 + 37 * i for i in range(40)], False),  # most requests cut in two, some whole
    (3, 128, 2, [3000, 130, 1500], True),    # M = 256: two q blocks per request, causal border
    (5, 96, 1, [1, 64, 2000, 700, 5000], False),  # half-filled second CTA, a one-token request
    # at least one request per scheduler part: everything is merged in the kernel, however many pieces (here one long
    # request over dozens of parts among short ones); below that, pieces beyond three go to the combine kernel
    (80, 128, 1, [20000] + [100 + 3 * i for i in range(79)], False),
])
def test_mla_pair_in_kernel_merge_equals_combine(mods, b, h, s_q, seqlens, causal):
    """Requests cut over several scheduler parts: the pair kernel's own merge (the part that finishes last adds the
    others' fp32 partials in split order, mla_combine.cu:72-153) against the partials + combine-kernel path on the same
    inputs -- same arithmetic in the same order, so at most an ulp of the bf16 output apart; the counters are zero
    again afterwards (the buffer serves the next call as is), and the call is repeatable on it."""
    fm = mods[1]
    c = make_case(fm, b, h, s_q, seqlens, nan_pad=False)
    sl = c["sl"].to(DEV)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1)
    assert (np.diff(ns.cpu().numpy()) > 1).any(), "expected at least one split request"
    a = (c["qn"], c["qr"], c["kl"], c["kr"], c["qs"], c["ks"], c["bt"].to(DEV), sl, 512, meta, ns)
    ms = fm.new_merge_state(b, s_q, h, DEV)
    o_c, l_c = fm.flash_mla_ckv_fp8_per_token(*a, causal=causal, merge_state=False)
    for _ in range(3):
        o_m, l_m = fm.flash_mla_ckv_fp8_per_token(*a, causal=causal, merge_state=ms)
        torch.cuda.synchronize()
        assert int(ms.abs().max()) == 0, "merge counters not back at zero"
        assert torch.equal(l_m, l_c) or float((l_m - l_c).abs().max()) < 1e-6
        d = (o_m.float() - o_c.float()).abs()
        assert float((d / o_c.float().abs().clamp_min(1e-3)).max()) < 2 ** -7, float(d.max())
        assert float((d > 0).float().mean()) < 1e-3  # and nearly everywhere the same bits


def test_mla_pair_is_timing_independent(tmp_path):
    """The CTA-pair kernel's hand-offs (row-reference moves, tile order, P / S / V rings, cluster relays, in-kernel
    merge) must not depend on who is fast: the diagnostic instantiation with a pseudo-random 0-4 us delay at each of its
    26 trace sites (B200_MLA_TRACE=5: warps, softmax groups and the two CTAs drift by whole tiles) has to reproduce the
    production instantiation's output bit for bit, run after run, under both tile schedules.  The same for the
    instantiation that rescales O behind a moving tile's weighted pass (B200_MLA_DEFER_RESCALE=1): undisturbed and
    disturbed it has to give the default instantiation's bits (the arithmetic is the same, only its place in time
    differs)."""
    import subprocess
    import sys
    helper = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chaos_case.py")
    outs = {}
    runs = (("plain", {}, 1), ("chaos", {"B200_MLA_TRACE": "5"}, 10), ("defer", {"B200_MLA_DEFER_RESCALE": "1"}, 1),
            ("defer_chaos", {"B200_MLA_DEFER_RESCALE": "1", "B200_MLA_TRACE": "5"}, 8))
    for name, env_extra, repeats in runs:
        env = dict(os.environ)
        for k in ("B200_MLA_TRACE", "B200_MLA_DEFER_RESCALE", "B200_MLA_CHAOS_SITES"):
            env.pop(k, None)
        env.update(env_extra)
        path = str(tmp_path / (name + ".npz"))
        r = subprocess.run([sys.executable, helper, path, str(repeats)], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-3000:]
        outs[name] = np.load(path)
    plain = outs["plain"]
    n = 0
    for name in ("chaos", "defer", "defer_chaos"):
        for key in outs[name].files:
            base = key.rsplit("_", 1)[0] + "_0"
            assert np.array_equal(outs[name][key], plain[base]), "%s (%s) differs from the undisturbed run" % (key, name)
            n += 1
    assert n >= 7 * 2 * (10 + 1 + 8) * 2


@pytest.mark.parametrize("h,s_q", [(128, 1), (16, 1)])
def test_mla_reference_moves_under_growing_logits(mods, h, s_q):
    """Keys whose norm grows along the sequence make the row maximum (and k_scale) climb tile after tile, so the lazy
    (m_ref, p_scale) reference has to move again and again: exercises the slow path, the O rescale in TMEM and, in
    the pair kernel, the version / hand-off protocol between the two softmax groups."""
    fm = mods[1]
    torch.manual_seed(5)
    b, seq = 2, 1536
    c = make_case(fm, b, h, s_q, [seq, seq - 200], nan_pad=False)
    n_pages = c["kl"].shape[0]
    kv = torch.randn(n_pages * 64, 1, 576, device=DEV)
    # token t of request i lives at page bt[i, t // 64], row t % 64: scale it by 2^(t / 96) (+16 octaves over the sequence)
    scale = torch.ones(n_pages, 64, device=DEV)
    for i in range(b):
        t = torch.arange(seq, device=DEV)
        pages = c["bt"][i].to(DEV).long()[t // 64]
        scale[pages, t % 64] = torch.exp2(t.float() / 96.0)
    kv = (kv * scale.view(-1, 1, 1) * 0.05).to(torch.bfloat16)
    fm.quantize_and_cache_k(kv, c["kl"].view(-1, 1, 512), c["ks"].view(-1, 1, 1), c["kr"].view(-1, 1, 64),
                            torch.arange(n_pages * 64, dtype=torch.int32, device=DEV), 512)
    out, lse, args, meta, ns = run_mla(fm, c, h, s_q, False)
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args)
    assert np.isfinite(out).all() and np.isfinite(lse).all()
    assert np.allclose(lse, sl, rtol=1e-3, atol=2e-3)
    assert rel(out, so) < 5e-2
    # twice the same launch: bit-identical (the recompute-on-stale-reference path must not depend on timing)
    out2, lse2, _, _, _ = run_mla(fm, c, h, s_q, False)
    assert np.array_equal(out, out2) and np.array_equal(lse, lse2)


def test_mla_pair_full_size_properties(mods):
    """Metric shape (b = 128, seq = 4096, h = 128): page-permutation invariance, bit for bit, and lse against torch on
    the de-quantised cache for a sample of requests."""
    fm = mods[1]
    b, h, seq = 128, 128, 4096
    c = make_case(fm, b, h, 1, [seq] * b, nan_pad=False)
    out1, lse1, _, _, _ = run_mla(fm, c, h, 1, False)
    n_pages = c["kl"].shape[0]
    perm = torch.randperm(n_pages - 1) + 1
    inv = torch.zeros(n_pages, dtype=torch.int64)
    inv[perm] = torch.arange(1, n_pages)
    c2 = dict(c)
    src = torch.cat([torch.zeros(1, dtype=torch.int64), inv[1:]]).to(DEV)
    c2["kl"], c2["kr"], c2["ks"] = c["kl"][src].contiguous(), c["kr"][src].contiguous(), c["ks"][src].contiguous()
    c2["bt"] = perm[c["bt"].long() - 1].to(torch.int32)
    out2, lse2, _, _, _ = run_mla(fm, c2, h, 1, False)
    assert np.array_equal(out1, out2) and np.array_equal(lse1, lse2)
    for i in (0, 77):
        pages = c["bt"][i, : seq // 64].long().to(DEV)
        kn = c["kl"][pages].view(-1, 512).view(torch.float8_e4m3fn).float() * c["ks"][pages].view(-1, 1)
        kr = c["kr"][pages].view(-1, 64).float() * c["ks"][pages].view(-1, 1)
        qn = c["qn"][i, 0].view(torch.float8_e4m3fn).float() * c["qs"][i, 0]
        qr = c["qr"][i, 0].float() * c["qs"][i, 0]
        logits = (qn @ kn.t() + qr @ kr.t()) * 576 ** -0.5
        assert np.allclose(lse1[i, :, 0], torch.logsumexp(logits, dim=-1).cpu().numpy(), rtol=1e-3, atol=1e-3)
        assert rel(out1[i, 0], (torch.softmax(logits, -1) @ kn).cpu().numpy()) < 5e-2


FULL_SIZE = [
    ("cfg2-v2lite", 128, 4096, 16, 1, (0, 41, 127)),
    ("metric-v3", 128, 4096, 128, 1, (0, 77)),
    ("cfg4-tp8-per-gpu", 256, 8192, 16, 1, (3, 255)),
    ("cfg5-mtp", 64, 4096, 16, 4, (0, 63)),
]


@pytest.mark.parametrize("name,b,seq,h,s_q,sample", FULL_SIZE, ids=[c[0] for c in FULL_SIZE])
def test_mla_full_size_vs_faithful_oracle(mods, name, b, seq, h, s_q, sample):
    """BASELINE.json shapes at full size: sampled requests against the reference-faithful recurrence (with the real
    scheduler metadata, so split requests go through the fp32 partials and the combine) at north_star's 1e-2."""
    fm = mods[1]
    sl = [seq] * b if s_q == 1 else [seq - 7 * (i % 5) for i in range(b)]
    c = make_case(fm, b, h, s_q, sl, nan_pad=False)
    out, lse, args, meta, ns = run_mla(fm, c, h, s_q, s_q > 1)
    fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta, ns, causal=s_q > 1, only_requests=set(sample))
    for i in sample:
        assert rel(out[i], fo[i]) < 1e-2, (i, rel(out[i], fo[i]))
        assert np.allclose(lse[i], fl[i], rtol=1e-3, atol=1e-3)


def test_mla_combine_direct_vs_oracle(mods):
    """a-5 on its own: b200_mla_combine through the C ABI against oracle.combine_partials (itself within one bf16 ulp of
    the reference kernel's source run on CPU): requests with 1 (untouched), 2, 3 and 7 splits, -inf partials (a split
    whose rows saw no token), an all -inf row (lse = +inf, out = 0); <= 1 bf16 ulp on out, lse to 1e-6."""
    L = mods[0].lib()
    rng = np.random.default_rng(7)
    h, s_q, dv = 16, 2, 512
    M = h * s_q
    n_splits = [1, 2, 3, 7, 1, 2]
    b = len(n_splits)
    ns = np.concatenate([[0], np.cumsum(n_splits)]).astype(np.int32)
    total = int(ns[-1])
    o_acc = rng.standard_normal((total, M, dv)).astype(np.float32)
    lse_acc = (rng.standard_normal((total, M)) * 3).astype(np.float32)
    lse_acc[ns[2] + 1, :5] = -np.inf          # one split of request 2 empty for 5 rows
    lse_acc[ns[5]:ns[5] + 2, 3] = -np.inf     # a row of request 5 with no token in any split
    out0 = rng.standard_normal((b, M, dv)).astype(np.float32)
    out0 = ofp8.round_to_bf16(out0)
    lse0 = rng.standard_normal((b, M)).astype(np.float32)
    want_out, want_lse = out0.copy(), lse0.copy()
    omla.combine_partials(o_acc, lse_acc, ns, want_out, want_lse)
    d_o, d_l, d_ns = torch.from_numpy(o_acc).to(DEV), torch.from_numpy(lse_acc).to(DEV), torch.from_numpy(ns).to(DEV)
    out = torch.from_numpy(out0).to(torch.bfloat16).to(DEV)                    # [b, s_q * h, dv]
    lse = torch.from_numpy(lse0).to(DEV)                                       # flat [b, M]; (h, s_q) permutation below
    # the kernel writes lse as [b, h, s_q] (flash_api.cpp:392-394); row m of the partials is (s, hh) = (m // h, m % h)
    lse_hs = lse.view(b, s_q, h).transpose(1, 2).contiguous()
    mods[0].check(L.b200_mla_combine(d_o.data_ptr(), d_l.data_ptr(), d_ns.data_ptr(), out.data_ptr(), lse_hs.data_ptr(),
                                     b, s_q, h, 0, mods[0].current_stream()), "combine")
    torch.cuda.synchronize()
    got_out = out.float().cpu().numpy()
    got_lse = lse_hs.transpose(1, 2).contiguous().view(b, M).cpu().numpy()
    for i, n in enumerate(n_splits):
        if n == 1:                               # mla_combine.cu:31-33: single-split requests are not touched
            assert np.array_equal(got_out[i], out0[i]) and np.array_equal(got_lse[i], lse0[i])
    ulp = np.abs(got_out - want_out) / np.maximum(np.abs(want_out), 1e-30)
    assert (ulp <= 2.0 ** -7).all(), ulp.max()   # one bf16 ulp
    assert np.isposinf(got_lse[5, 3]) and np.isposinf(want_lse[5, 3]) and (got_out[5, 3] == 0).all()
    fin = np.isfinite(want_lse)
    assert np.array_equal(fin, np.isfinite(got_lse)) and np.allclose(got_lse[fin], want_lse[fin], rtol=1e-6, atol=1e-6)


def test_metadata_idle_batch(mods):
    """batch_size 0 (an IDLE data-parallel attention rank, flash_api.cpp:27-63 accepts it): num_splits = [0] and no part
    has work."""
    fm = mods[1]
    meta, ns = fm.get_mla_metadata(torch.empty(0, dtype=torch.int32, device=DEV), 16, 1)
    assert ns.cpu().tolist() == [0] and meta.shape[1] == 8
    assert (meta[:, 0] >= 0).all() and (meta[:, 2] == -1).all()


@pytest.mark.parametrize("h,s_q", [(64, 1), (16, 4), (48, 1)])
def test_mla_m_le_64_default_routing(mods, h, s_q):
    """32 < M <= 64 fills one CTA of the pair kernel by default (faster than the single-CTA SwapAB kernel at HB = 64)."""
    fm = mods[1]
    c = make_case(fm, 3, h, s_q, [700, 129, 65])
    out, lse, args, meta, ns = run_mla(fm, c, h, s_q, s_q > 1)
    fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta, ns, causal=s_q > 1)
    assert rel(out, fo) < 1e-2 and np.allclose(lse, fl, rtol=1e-3, atol=1e-3)


def test_mla_decode_low_noise_input_is_tight(mods):
    """With a constant-sign V the e4m3 rounding of P averages out: the kernel must agree with fp32 attention to well
    under 1e-2 -- isolates real bugs from quantisation noise."""
    fm = mods[1]
    torch.manual_seed(3)
    b, h, s_q = 2, 16, 1
    c = make_case(fm, b, h, s_q, [900, 400], nan_pad=False)
    # rebuild the pool from an all-positive latent so sum(P V) does not cancel
    n_pages = c["kl"].shape[0]
    kv = torch.rand(n_pages * 64, 1, 576, device=DEV).add_(0.5).to(torch.bfloat16)
    kl, ks, kr = c["kl"].view(-1, 1, 512), c["ks"].view(-1, 1, 1), c["kr"].view(-1, 1, 64)
    fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=DEV), 512)
    out, lse, args, meta, ns = run_mla(fm, c, h, s_q, False)
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args)
    assert rel(out, so) < 6e-3
    assert np.allclose(lse, sl, rtol=1e-3, atol=1e-3)


def test_mla_decode_full_size_properties(mods):
    """BASELINE shape (b=128, seq=4096, h=16): size-independent properties instead of a CPU oracle --
    (1) result is invariant to a permutation of the physical pages (block-table gather),
    (2) appending fully masked garbage pages changes nothing, (3) lse equals a bf16 torch attention on the
    de-quantised cache for a sample of requests."""
    fm = mods[1]
    b, h, seq = 128, 16, 4096
    c = make_case(fm, b, h, 1, [seq] * b, nan_pad=False)
    out1, lse1, _, _, _ = run_mla(fm, c, h, 1, False)
    # (1) move every page to a new physical location
    n_pages = c["kl"].shape[0]
    perm = torch.randperm(n_pages - 1) + 1
    inv = torch.zeros(n_pages, dtype=torch.int64)
    inv[perm] = torch.arange(1, n_pages)
    c2 = dict(c)
    src = torch.cat([torch.zeros(1, dtype=torch.int64), inv[1:]]).to(DEV)  # new page p holds old page inv[p]
    c2["kl"], c2["kr"], c2["ks"] = c["kl"][src].contiguous(), c["kr"][src].contiguous(), c["ks"][src].contiguous()
    c2["bt"] = perm[c["bt"].long() - 1].to(torch.int32)
    out2, lse2, _, _, _ = run_mla(fm, c2, h, 1, False)
    assert np.array_equal(out1, out2) and np.array_equal(lse1, lse2)
    # (3) lse vs torch on de-quantised data, 4 requests
    for i in (0, 41, 127):
        pages = c["bt"][i, : seq // 64].long().to(DEV)
        kn = c["kl"][pages].view(-1, 512).view(torch.float8_e4m3fn).float() * c["ks"][pages].view(-1, 1)
        kr = c["kr"][pages].view(-1, 64).float() * c["ks"][pages].view(-1, 1)
        qn = c["qn"][i, 0].view(torch.float8_e4m3fn).float() * c["qs"][i, 0]
        qr = c["qr"][i, 0].float() * c["qs"][i, 0]
        logits = (qn @ kn.t() + qr @ kr.t()) * 576 ** -0.5
        ref_lse = torch.logsumexp(logits, dim=-1).cpu().numpy()
        assert np.allclose(lse1[i, :, 0], ref_lse, rtol=1e-3, atol=1e-3)
        ref_out = (torch.softmax(logits, -1) @ kn).cpu().numpy()
        assert rel(out1[i, 0], ref_out) < 5e-2


def test_block_table_gather_bit_exact():
    """a-10: block_table = req_to_page[req_pool_indices] -- integer gather done by torch on device must equal the
    allocator oracle bit for bit (reference: flashmla_backend.py:259)."""
    alloc = opool.KVAllocator(size=64 * 64, max_batch_size=8, max_context_len=1024)
    lens = [0] * 8
    for req, need in [(3, 200), (1, 65), (3, 1), (7, 1000), (1, 64)]:
        alloc.alloc(req, need, lens[req])
        lens[req] += need
    idx = torch.tensor([7, 1, 3], device=DEV)
    got = torch.from_numpy(alloc.req_to_page).to(DEV)[idx].cpu().numpy()
    assert np.array_equal(got, opool.block_table_gather(alloc.req_to_page, [7, 1, 3]))
