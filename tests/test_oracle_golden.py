"""CPU: the oracle restatements against fixtures generated from the reference's own Python
(tests/golden/make_golden.py).  These pin the oracle; the GPU parity tests then compare CUDA against the oracle."""
import os

import numpy as np
import pytest

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from oracle import mla as omla
from oracle import pool as opool


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_e4m3_codec_matches_torch():
    torch = pytest.importorskip("torch")
    codes = np.arange(256, dtype=np.uint8)
    ours = ofp8.e4m3_to_f32(codes)
    ref = torch.from_numpy(codes).view(torch.float8_e4m3fn).float().numpy()
    assert np.array_equal(np.isnan(ours), np.isnan(ref))
    assert np.array_equal(ours[~np.isnan(ours)], ref[~np.isnan(ref)])
    x = (torch.randn(20000) * 100).clamp(-440, 440)
    x = torch.cat([x, torch.tensor([0.0, -0.0, 1e-9, 447.9, -448.0, 0.0009765625, 0.00146484375])])
    got = ofp8.f32_to_e4m3(x.numpy())
    want = x.to(torch.float8_e4m3fn).view(torch.uint8).numpy()
    assert np.array_equal(got, want)
    # saturation (cvt.rn.satfinite) where torch would give NaN
    assert ofp8.f32_to_e4m3(np.array([1e6, -1e6], np.float32)).tolist() == [0x7E, 0xFE]


def test_bf16_rounding_matches_torch():
    torch = pytest.importorskip("torch")
    x = torch.randn(50000) * 37
    assert np.array_equal(ofp8.round_to_bf16(x.numpy()), x.to(torch.bfloat16).float().numpy())


def test_pool_fallback_quantiser_golden(golden_dir):
    g = _load(golden_dir, "pool_per_token_head.npz")
    codes, scale, rope = opool.pool_fallback_quantize(g["key"][:, 0, :])
    loc = g["loc"]
    assert np.array_equal(codes, g["k_lora"][loc, 0, :])
    assert np.array_equal(scale.view(np.uint32), g["k_scale"][loc, 0, :].view(np.uint32))
    assert np.array_equal(rope, g["k_rope"][loc, 0, :])
    deq_l, deq_r = opool.pool_fallback_dequantize(g["k_lora"], g["k_scale"], g["k_rope"], loc)
    assert np.array_equal(deq_l, g["deq_lora"])
    assert np.array_equal(deq_r, g["deq_rope"])


def test_cuda_quantiser_oracle_close_to_reference_fallback(golden_dir):
    """The CUDA quantiser restatement (scale + 1e-12, multiply by the reciprocal) vs the reference's own torch
    fall-back: same scales to 1e-6 relative, e4m3 codes equal except rare one-step rounding flips."""
    g = _load(golden_dir, "pool_per_token_head.npz")
    key = g["key"][:, 0, :]
    codes, scale, rope = omla.quantize_rows(key)
    loc = g["loc"]
    ref_scale = g["k_scale"][loc, 0, 0]
    nz = ref_scale > 1e-20
    assert np.allclose(scale[nz], ref_scale[nz], rtol=1e-6)
    ref_codes = g["k_lora"][loc, 0, :]
    diff = codes[nz] != ref_codes[nz]
    assert diff.mean() < 5e-3
    a, b = ofp8.e4m3_to_f32(codes[nz][diff]), ofp8.e4m3_to_f32(ref_codes[nz][diff])
    assert np.all(np.abs(a - b) <= 0.126 * np.maximum(np.abs(a), np.abs(b)) + 2 ** -9)
    # dequantised round trip within e4m3 resolution (test_cache_kernels.py:59-66 uses rel-L1 < 1e-3 vs torch)
    deq = ofp8.e4m3_to_f32(codes) * scale[:, None]
    rel_l1 = np.abs(deq - key[:, :512]).sum() / np.abs(key[:, :512]).sum()
    assert rel_l1 < 3e-2


def test_indexed_dequant_oracle_matches_fallback(golden_dir):
    g = _load(golden_dir, "pool_per_token_head.npz")
    lora, rope = omla.dequantize_ckv_fused_indexed(g["k_lora"], g["k_rope"], g["k_scale"], g["loc"])
    assert np.array_equal(lora, g["deq_lora"])
    assert np.array_equal(rope, g["deq_rope"])


def test_allocator_golden(golden_dir):
    g = _load(golden_dir, "allocator.npz")
    alloc = opool.KVAllocator(size=64 * 40, max_batch_size=4, max_context_len=640, page_size=64)
    lens = [0, 0, 0]
    locs = []
    for req, need in zip(g["reqs"], g["needs"]):
        locs.append(alloc.alloc(int(req), int(need), lens[req]))
        lens[req] += int(need)
    assert np.array_equal(np.concatenate(locs), g["kv_locs"])
    assert np.array_equal(alloc.req_to_page, g["req_to_page"])
    bt = opool.block_table_gather(alloc.req_to_page, [2, 0])
    assert np.array_equal(bt, g["req_to_page"][[2, 0]])


def test_native_sdpa_golden(golden_dir):
    g = _load(golden_dir, "native_sdpa_decode.npz")
    out = opool.native_sdpa_decode(g["q"], g["k_cache"], g["req_to_token"], g["seq_lens"])
    ref = g["out"]
    assert np.linalg.norm(out - ref) / np.linalg.norm(ref) < 1e-2  # reference computes in bf16


def test_block_fp8_golden(golden_dir):
    g = _load(golden_dir, "block_fp8.npz")
    q, s = ogemm.per_token_group_quant_fp8(g["x"])
    assert np.array_equal(q, g["xq"])
    # all-zero groups: the torch test oracle clamps amax to eps in the INPUT dtype (bf16(1e-10) = 1.0004e-10), the CUDA
    # kernel (sgl_fp8_quantization.cu:372) and this restatement use the fp32 eps; everything else is bit-equal.
    zero_group = np.abs(g["x"]).reshape(s.shape[0], s.shape[1], 128).max(-1) < 1e-10
    assert np.array_equal(s.view(np.uint32)[~zero_group], g["xs"].view(np.uint32)[~zero_group])
    assert np.allclose(s[zero_group], g["xs"][zero_group], rtol=1e-3)
    c = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(g["xq"]), g["xs"], ofp8.e4m3_to_f32(g["wq"]), g["ws"])
    assert np.allclose(c, g["c"], rtol=1e-5, atol=1e-4)


def test_test_side_quantisers_match_the_references(golden_dir):
    """tests/helpers.py stands in for per_token_cast_to_fp8 / per_block_cast_to_fp8 of the reference's GEMM tests
    (3rdparty/deep_gemm/tests/test_core.py:9-25): bit-equal codes and scales on the golden inputs."""
    import torch
    from helpers import quantize_lhs_1x128, quantize_rhs_128x128
    g = _load(golden_dir, "block_fp8.npz")
    xq, xs = quantize_lhs_1x128(torch.from_numpy(g["x"]).to(torch.bfloat16))
    assert np.array_equal(xq.view(torch.uint8).numpy(), g["xq_cast"])
    assert np.array_equal(xs.numpy().view(np.uint32), g["xs_cast"].view(np.uint32))
    wq, ws = quantize_rhs_128x128(torch.from_numpy(g["w"]).to(torch.bfloat16))
    assert np.array_equal(wq.view(torch.uint8).numpy(), g["wq"])
    assert np.array_equal(ws.numpy().view(np.uint32), g["ws"].view(np.uint32))


def test_semantic_vs_faithful_attention_agree():
    """The two attention restatements (plain fp64 softmax vs the reference's page-wise FP8-P recurrence) agree
    within the reference's own criterion (tests/test_flash_mla_ckv_fp8_per_token.py:32-39, rel < 5e-2)."""
    rng = np.random.default_rng(0)
    b, s_q, h = 2, 1, 4
    seqlens = np.array([150, 64], np.int32)
    n_pages, d = 8, 576
    kv = ofp8.round_to_bf16(rng.standard_normal((n_pages * 64, 576)).astype(np.float32))
    kl = np.zeros((n_pages * 64, 1, 512), np.uint8)
    ks = np.zeros((n_pages * 64, 1, 1), np.float32)
    kr = np.zeros((n_pages * 64, 1, 64), np.float32)
    omla.quantize_and_cache_k(kv[:, None, :], kl, ks, kr, np.arange(n_pages * 64))
    q = ofp8.round_to_bf16(rng.standard_normal((b, s_q, h, 576)).astype(np.float32))
    qn, qs, qr = omla.quantize_ckv_per_token_head(q)
    bt = np.array([[3, 1, 5, 0], [2, 7, 0, 0]], np.int32)
    meta, ns = omla.get_mla_metadata(seqlens, s_q * h, 1, 8)
    args = (qn, qr, kl.reshape(n_pages, 64, 1, 512), kr.reshape(n_pages, 64, 1, 64), qs, ks.reshape(n_pages, 64, 1, 1),
            bt, seqlens)
    fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta, ns)
    so, sl = omla.flash_mla_ckv_fp8_semantic(*args)
    assert np.linalg.norm(fo - so) / np.linalg.norm(so) < 5e-2
    assert np.allclose(fl, sl, rtol=1e-4, atol=1e-4)


def test_metadata_invariants():
    """Scheduler output covers every block exactly once and num_splits is its prefix sum."""
    rng = np.random.default_rng(1)
    for sm in (8, 74, 148):
        seqlens = rng.integers(1, 9000, size=37).astype(np.int32)
        meta, ns = omla.get_mla_metadata(seqlens, 16, 1, sm)
        covered = np.zeros(len(seqlens), np.int64)
        for part in meta:
            b0, s0, b1, s1 = (int(v) for v in part[:4])
            if b0 >= len(seqlens):
                continue
            for bi in range(b0, b1 + 1):
                lo = s0 // 64 if bi == b0 else 0
                hi = -(-s1 // 64) if bi == b1 else -(-int(seqlens[bi]) // 64)
                covered[bi] += hi - lo
        assert np.array_equal(covered, -(-seqlens.astype(np.int64) // 64))
        assert ns[0] == 0 and np.all(np.diff(ns) >= 1)


def test_rope_oracle_matches_reference_forward_native(golden_dir):
    """oracle/rope.py::apply_rope vs DeepseekScalingRotaryEmbedding.forward_native of the reference
    (python/sglang/srt/layers/rotary_embedding.py:804-846; V3 yarn cache, both pairings): bit-exact bf16 outputs."""
    from oracle import rope as orope
    g = _load(golden_dir, "rope.npz")
    assert tuple(g["cache_shape_gptj"]) == (163840, 64)          # [max_position * factor, rotary_dim] = cos | sin
    for name, neox in (("gptj", False), ("neox", True)):
        rows = g["cos_sin_rows_" + name]                          # the cache rows of the vectors' positions
        ids = np.arange(len(rows))
        assert np.array_equal(orope.apply_rope(g["q_pe"], ids, rows, neox), g["q_out_" + name])
        assert np.array_equal(orope.apply_rope(g["k_pe"], ids, rows, neox), g["k_out_" + name])
    assert not np.array_equal(g["q_out_gptj"], g["q_out_neox"])  # the two pairings are different rotations


def test_fused_prologue_oracle_is_the_unfused_sequence(golden_dir):
    """rope_quantize_fused == rope -> quantize_ckv_per_token_head / quantize_and_cache_k, and a rotation preserves the
    pair norms (size-independent property), so scales computed from the latent only are untouched by it."""
    from oracle import rope as orope
    g = _load(golden_dir, "rope.npz")
    rng = np.random.default_rng(5)
    t, h = g["q_pe"].shape[:2]
    q_nope = ofp8.round_to_bf16(rng.standard_normal((t, h, 512)).astype(np.float32))
    kv = ofp8.round_to_bf16(rng.standard_normal((t, 576)).astype(np.float32))
    kv[:, 512:] = g["k_pe"][:, 0]
    rows, ids = g["cos_sin_rows_gptj"], np.arange(t)
    kl, ks, kr = np.zeros((64, 1, 512), np.uint8), np.zeros((64, 1, 1), np.float32), np.zeros((64, 1, 64), np.float32)
    loc = rng.permutation(64)[:t]
    n8, sc, qr = orope.rope_quantize_fused(q_nope, g["q_pe"], kv, ids, rows, False, loc, kl, ks, kr)
    q = np.concatenate([q_nope, g["q_out_gptj"]], -1)
    n8_u, sc_u, qr_u = omla.quantize_ckv_per_token_head(q)
    assert np.array_equal(n8, n8_u) and np.array_equal(sc, sc_u) and np.array_equal(qr, qr_u)
    deq_l, deq_r = omla.dequantize_ckv_fused_indexed(kl, kr, ks, loc)
    assert np.allclose(deq_r[:, 0], g["k_out_gptj"][:, 0], rtol=2 ** -7, atol=1e-6)
    x, y = g["q_pe"], g["q_out_gptj"]
    assert np.allclose(x[..., 0::2] ** 2 + x[..., 1::2] ** 2, y[..., 0::2] ** 2 + y[..., 1::2] ** 2, rtol=3e-2, atol=1e-3)


def test_pool_disagg_descriptors_and_moves_match_reference(golden_dir):
    """The host mirror of MLATokenToKVPool (sglang_fluentllm_b200/flashmla_backend.py) against the reference class run on
    CPU (mem_cache/memory_pool.py:747-806): transfer descriptors of both layouts (order, byte lengths, page item
    lengths, per-layer offsets) and move_kv_cache on the three-tensor layout, bit for bit."""
    torch = pytest.importorskip("torch")
    from sglang_fluentllm_b200 import flashmla_backend as fb
    g = _load(golden_dir, "pool_disagg.npz")
    for qm in ("per_token_head", "none"):
        pool = fb.MLATokenToKVPool(size=64, model_dtype=torch.bfloat16, kv_lora_rank=512, qk_rope_head_dim=64,
                                   layer_num=3, device="cpu", quant_method=qm)
        ptrs, lens, item_lens = pool.get_contiguous_buf_infos()
        bufs = [t for kv in pool.kv_buffer for t in (kv if isinstance(kv, tuple) else (kv,))]
        order = [[b.data_ptr() for b in bufs].index(p) for p in ptrs]
        assert order == g["order_" + qm].tolist()
        assert lens == g["lens_" + qm].tolist() and item_lens == g["item_lens_" + qm].tolist()
        assert pool.get_layerwise_buf_info_offsets(5) == g["offsets_" + qm].tolist()
        if int(g["size_bytes_" + qm]) >= 0:   # the reference's own method raises on the tuple layout (:779-784)
            assert pool.get_kv_size_bytes() == int(g["size_bytes_" + qm])
        else:
            assert pool.get_kv_size_bytes() == sum(lens)
    pool = fb.MLATokenToKVPool(size=64, model_dtype=torch.bfloat16, kv_lora_rank=512, qk_rope_head_dim=64,
                               layer_num=3, device="cpu", quant_method="per_token_head")
    for l, kv in enumerate(pool.kv_buffer):
        kv[0].copy_(torch.from_numpy(g["before_lora_%d" % l]))
        kv[1].copy_(torch.from_numpy(g["before_scale_%d" % l]))
        kv[2].copy_(torch.from_numpy(g["before_rope_%d" % l]).to(torch.bfloat16))
    pool.move_kv_cache(torch.from_numpy(g["move_tgt"]), torch.from_numpy(g["move_src"]))
    for l, kv in enumerate(pool.kv_buffer):
        assert np.array_equal(kv[0].numpy(), g["after_lora_%d" % l])
        assert np.array_equal(kv[1].numpy(), g["after_scale_%d" % l])
        assert np.array_equal(kv[2].float().numpy(), g["after_rope_%d" % l])
    # a moved token de-quantises to the same values (codes, scale and pre-divided rope travel together)
    src, tgt = g["move_src"], g["move_tgt"]
    a = opool.pool_fallback_dequantize(g["after_lora_0"], g["after_scale_0"], g["after_rope_0"], tgt[[0, 1, 3]])
    b = opool.pool_fallback_dequantize(g["before_lora_0"], g["before_scale_0"], g["before_rope_0"], src[[0, 1, 3]])
    for x, y in zip(a, b):
        assert np.array_equal(x, y, equal_nan=True)


def test_attention_oracles_pass_the_references_own_test(golden_dir):
    """The reference pins its kernel by one test: per-token-quantise random q / paged K (NaN-poisoned padding) and demand
    relative_L1 < 5e-2 against fp32 attention on the un-quantised tensors
    (3rdparty/flashmla-fp8/tests/test_flash_mla_ckv_fp8_per_token.py:33-40, 48-116).  The fixture holds that test's
    inputs and its `ref_mla` outputs, produced by the reference's own code on CPU.  Here every attention oracle the GPU
    parity tests check the CUDA kernels against has to pass that same test: the bf16 restatement (a-2) must *be*
    ref_mla, the semantic and the faithful FP8 restatements (a-1) must meet the reference's criterion."""
    torch = pytest.importorskip("torch")
    g = _load(golden_dir, "mla_test_reference.npz")

    def rel_l1(x, y):  # cal_diff's "relative_L1" (:38)
        return float(np.linalg.norm(x.astype(np.float64) - y) / np.linalg.norm(y.astype(np.float64)))

    def bf16(name):
        return torch.from_numpy(g[name]).view(torch.bfloat16)

    for tag in ("decode", "verify"):
        q, blocked_k = bf16(tag + "_q_bf16bits"), bf16(tag + "_blocked_k_bf16bits")
        seqlens, block_table, causal = g[tag + "_cache_seqlens"], g[tag + "_block_table"], bool(g[tag + "_causal"])
        want_o, want_lse = g[tag + "_out"], g[tag + "_lse"]
        b, s_q, h_q, _ = q.shape
        # a-2: bf16 Q / bf16 paged K -- the same arithmetic as ref_mla up to fp32-vs-fp64 and the bf16 rounding of out
        o, lse = omla.flash_mla_with_kvcache(q.float().numpy(), blocked_k.float().numpy(), block_table, seqlens, causal=causal)
        assert rel_l1(o, want_o) < 4e-3 and np.allclose(lse, want_lse, rtol=1e-4, atol=1e-4)
        # the test-side quantisation (:94-101), restated
        q_scale = q[..., :512].float().abs().amax(dim=-1, keepdim=True) / 448.0
        q_rope = (q[..., 512:].float() / q_scale).to(torch.bfloat16)
        q8 = (q[..., :512].float() / q_scale).to(torch.float8_e4m3fn).view(torch.uint8)
        k_scale = blocked_k[..., :512].float().abs().amax(dim=-1, keepdim=True) / 448.0
        k_rope = (blocked_k[..., 512:].float() / k_scale).to(torch.bfloat16)
        k8 = (blocked_k[..., :512].float() / k_scale).to(torch.float8_e4m3fn).view(torch.uint8)
        assert np.array_equal(q_scale.numpy(), g[tag + "_q_scale"])
        assert np.array_equal(k_scale.numpy(), g[tag + "_k_scale"], equal_nan=True)
        assert int(q8.long().sum()) == int(g[tag + "_q_nope_fp8_sum"]) and int(k8.long().sum()) == int(g[tag + "_k_lora_fp8_sum"])
        args = (q8.numpy(), q_rope.float().numpy(), k8.numpy(), k_rope.float().numpy(), q_scale.numpy(), k_scale.numpy(),
                block_table, seqlens)
        so, sl = omla.flash_mla_ckv_fp8_semantic(*args, causal=causal)
        assert rel_l1(so, want_o) < 5e-2 and rel_l1(sl, want_lse) < 5e-2
        assert not np.isnan(so).any(), "NaN leaked from the poisoned padding rows"
        meta, ns = omla.get_mla_metadata(seqlens, s_q * h_q, 1, 8)
        fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta, ns, causal=causal)
        assert rel_l1(fo, want_o) < 5e-2 and rel_l1(fl, want_lse) < 5e-2
        assert not np.isnan(fo).any()


def test_quantiser_oracles_against_the_references_own_test(golden_dir):
    """3rdparty/flashmla-fp8/tests/test_cache_kernels.py compares the CUDA quantisers with a torch reference on rows of
    varying range and offset (relative_L1 < 1e-3, :59-67, :88-146).  The fixture holds that test's inputs and torch
    reference outputs, produced by its own generator on CPU.  The restatements the CUDA kernels are bit-exact with
    (oracle/mla.py) give identical scales, identical rope values and identical cached-key scales/rope; the e4m3 codes
    differ from the torch reference on < 0.1 % of the elements, by one code, and only where x / scale sits on a rounding
    midpoint: the torch reference divides by the scale (:91-92), the CUDA kernel multiplies by a rounded reciprocal
    (csrc/kernels/ckv_fp8_per_token/cache_kernels.cu:79,84), so exact ties (common with bf16 inputs) break differently.
    That alone is relative_L1 = 2.5e-3 on this sample -- the reference kernel's own arithmetic does not meet the test's
    1e-3 on it either; stated here instead of hidden."""
    torch = pytest.importorskip("torch")
    g = _load(golden_dir, "cache_kernel_test_reference.npz")

    def rel_l1(x, y):
        x, y = np.asarray(x, np.float64), np.asarray(y, np.float64)
        return float(np.linalg.norm(x - y) / np.linalg.norm(y))

    def check_codes(codes, want, x, scale):
        got_v, want_v = ofp8.e4m3_to_f32(codes).astype(np.float64), ofp8.e4m3_to_f32(want).astype(np.float64)
        diff = codes != want
        assert diff.mean() < 1e-3 and rel_l1(got_v, want_v) < 5e-3
        # every differing element: adjacent codes, and the exact quotient lies (to fp32 resolution) on their midpoint
        exact = (x.astype(np.float64) / scale.astype(np.float64))[diff]
        mid = (got_v[diff] + want_v[diff]) / 2
        assert np.all(np.abs(codes[diff].astype(np.int16) - want[diff].astype(np.int16)) == 1)
        assert np.all(np.abs(exact - mid) <= np.abs(mid) * 2.0 ** -22)

    query = torch.from_numpy(g["query_bf16bits"]).view(torch.bfloat16).float().numpy()
    key = torch.from_numpy(g["key_bf16bits"]).view(torch.bfloat16).float().numpy()
    codes, scale, rope = omla.quantize_ckv_per_token_head(query)
    assert np.array_equal(scale, g["q_scale"]) and np.array_equal(rope, g["q_rope"])
    check_codes(codes, g["q_nope"], query[..., :512], scale)
    slots, loc = int(g["num_slots"]), g["loc"]
    kl, ks, kr = np.zeros((slots, 1, 512), np.uint8), np.zeros((slots, 1, 1), np.float32), np.zeros((slots, 1, 64), np.float32)
    omla.quantize_and_cache_k(key.reshape(-1, 1, 576), kl, ks, kr, loc)
    assert rel_l1(ks[loc], g["k_scale"]) < 1e-3 and rel_l1(kr[loc], g["k_rope"]) < 1e-3
    check_codes(kl[loc], g["k_lora"], key.reshape(-1, 1, 576)[..., :512], ks[loc])
    untouched = np.ones(slots, bool)
    untouched[loc] = False
    assert not kl[untouched].any() and not ks[untouched].any()   # only the addressed slots are written


def test_oracle_bit_exact_vs_reference_kernels_fixture(golden_dir):
    """The restatements the CUDA kernels are checked against (oracle/mla.py: get_mla_metadata, quantize_ckv_per_token_head,
    quantize_and_cache_k, dequantize_ckv_fused_indexed) against the outputs of the reference's OWN kernel sources
    (3rdparty/flashmla-fp8/csrc/kernels/get_mla_metadata.cu, ckv_fp8_per_token/cache_kernels.cu,
    3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu) compiled for the host and executed on CPU (oracle/ref_build/):
    every int32, e4m3 code, fp32 scale bit and bf16 bit identical."""
    g = _load(golden_dir, "ref_kernels_cpu.npz")
    for i in range(int(g["meta_cases"])):
        sm, hpk, b = (int(v) for v in g["meta_case_%d" % i])
        meta, ns = omla.get_mla_metadata(g["meta_seqlens_%d" % i], hpk, 1, sm)
        want = g["meta_out_%d" % i]
        assert meta.shape == want.shape
        assert np.array_equal(meta[:, :5], want[:, :5])          # columns 5-7 are never written by the reference kernel
        assert np.array_equal(ns, g["meta_num_splits_%d" % i])
    bits = lambda a: ofp8.f32_to_bf16_bits(a)
    q = ofp8.bf16_bits_to_f32(g["q_bits"])
    codes, scale, rope = omla.quantize_ckv_per_token_head(q)
    assert np.array_equal(codes, g["q_nope"]) and np.array_equal(scale.view(np.uint32), g["q_scale"].view(np.uint32))
    assert np.array_equal(bits(rope), g["q_rope_bits"])
    slots = int(g["k_slots"])
    kl, ks, kr = np.zeros((slots, 1, 512), np.uint8), np.zeros((slots, 1, 1), np.float32), np.zeros((slots, 1, 64), np.float32)
    omla.quantize_and_cache_k(ofp8.bf16_bits_to_f32(g["k_bits"]), kl, ks, kr, g["k_loc"])
    assert np.array_equal(kl, g["k_lora"]) and np.array_equal(ks.view(np.uint32), g["k_scale"].view(np.uint32))
    assert np.array_equal(bits(kr), g["k_rope_bits"])
    dl, dr = omla.dequantize_ckv_fused_indexed(kl, kr, ks, g["deq_idx"])
    assert np.array_equal(bits(dl), g["deq_lora_bits"]) and np.array_equal(bits(dr), g["deq_rope_bits"])
    # G-2: 1x128 activation quantiser (3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu:357-437)
    aq, a_s = ogemm.per_token_group_quant_fp8(ofp8.bf16_bits_to_f32(g["act_bits"]))
    assert np.array_equal(aq, g["act_q"]) and np.array_equal(a_s.view(np.uint32), g["act_scales"].view(np.uint32))
    # a-5: split-KV combine (mla_combine.cu:16-154).  exp2f / log2f and the order of the 32-lane sums differ between libm +
    # butterfly and numpy, so this one is pinned to fp32 rounding, not to the bit: lse within 2 ulp, out within one bf16 ulp
    # on a handful of elements; -inf partials and single-split requests handled identically.
    ns = g["comb_num_splits"]
    b, M = len(ns) - 1, g["comb_lse_accum"].shape[1]
    out, lse = np.zeros((b, M, 512), np.float32), np.zeros((b, M), np.float32)
    omla.combine_partials(g["comb_o_accum"], g["comb_lse_accum"], ns, out, lse)
    want_bits, want_lse = g["comb_out_bits"], g["comb_lse"]
    assert np.allclose(lse, want_lse, rtol=3e-7, atol=0) and not want_bits[1].any() and not out[1].any()
    got_bits = bits(out).astype(np.int32)
    ulp = np.abs(got_bits - want_bits.astype(np.int32))
    assert ulp.max() <= 1 and (ulp != 0).mean() < 5e-3


def test_oracle_bit_exact_vs_reference_kernels_live():
    """Same pin, live: random sweeps against oracle/_ref/libflashmla_ref_cpu.so when it has been built (it is built by
    __graft_entry__.build() wherever /root/reference exists, and travels to the GPU box)."""
    from oracle import ref_cpu
    if not ref_cpu.available():
        pytest.skip("oracle/_ref not built (no reference tree on this machine)")
    rng = np.random.default_rng(7)
    for sm, hpk in ((148, 16), (148, 128), (148, 256), (132, 64), (8, 16)):
        for b in (1, 2, 37, 128, 300):
            seqlens = rng.integers(1, 20000, size=b).astype(np.int32)
            if b == 2:
                seqlens[:] = 1
            want, want_ns = ref_cpu.get_mla_metadata(seqlens, hpk, 1, sm)
            meta, ns = omla.get_mla_metadata(seqlens, hpk, 1, sm)
            assert np.array_equal(meta[:, :5], want[:, :5]) and np.array_equal(ns, want_ns), (sm, hpk, b)
    x = (rng.standard_normal((96, 2, 576)) * rng.uniform(1e-4, 60, (96, 2, 1))).astype(np.float32)
    x[0] = 0
    xb = ofp8.f32_to_bf16_bits(x)
    c, sc, r = ref_cpu.quantize_ckv_per_token_head(xb)
    oc, osc, orr = omla.quantize_ckv_per_token_head(ofp8.bf16_bits_to_f32(xb))
    assert np.array_equal(c, oc) and np.array_equal(sc.view(np.uint32), osc.view(np.uint32))
    assert np.array_equal(r, ofp8.f32_to_bf16_bits(orr))
    for m, k in ((1, 128), (5, 512), (37, 1536), (3, 7168)):
        xa = (rng.standard_normal((m, k)) * rng.uniform(1e-3, 50, (m, 1))).astype(np.float32)
        ab = ofp8.f32_to_bf16_bits(xa)
        q_ref, s_ref = ref_cpu.per_token_group_quant_fp8(ab)
        q_o, s_o = ogemm.per_token_group_quant_fp8(ofp8.bf16_bits_to_f32(ab))
        assert np.array_equal(q_ref, q_o) and np.array_equal(s_ref.view(np.uint32), s_o.view(np.uint32))
