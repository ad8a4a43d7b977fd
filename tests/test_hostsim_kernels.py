"""CPU: the CUDA SOURCE of the product's streaming kernels (csrc/small_kernels.cu), compiled for the host and run on CPU
threads (tests/hostsim/), against the reference's own kernel source run the same way (oracle/_ref) and against the
oracle -- bit for bit.  This is a checker over the kernel arithmetic and indexing that needs no GPU; the `-m gpu` tests
through the C ABI remain the parity tests proper."""
import ctypes

import numpy as np
import pytest

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from oracle import mla as omla
from oracle import ref_cpu
from oracle import rope as orope

_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
bits = ofp8.f32_to_bf16_bits
vals = ofp8.bf16_bits_to_f32


@pytest.fixture(scope="module")
def host():
    from hostsim import build_hostsim
    lib = ctypes.CDLL(build_hostsim.build())
    P, I, L, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float
    for name, args in {
        "host_mla_get_metadata": [P, I, I, P, P],
        "host_quantize_q": [P, L, P, P, P],
        "host_quantize_and_cache_k": [P, L, L, P, P, P, P],
        "host_dequantize_indexed": [P, P, P, P, L, L, P, P],
        "host_mla_combine": [P, P, P, P, P, I, I, I],
        "host_act_quant_1x128_strided": [P, L, L, L, L, P, P, L, F],
        "host_rope_quantize_fused": [P, L, L, P, L, L, P, L, P, P, I, L, I, P, P, P, P, P, P, P],
    }.items():
        getattr(lib, name).argtypes = args
        getattr(lib, name).restype = I
    return lib


def aligned(shape, dtype):
    return ref_cpu._aligned(shape, dtype)


def bf16_rows(rng, shape, lo=1e-3, hi=40.0):
    x = (rng.standard_normal(shape) * rng.uniform(lo, hi, shape[:-1] + (1,))).astype(np.float32)
    out = aligned(shape, np.uint16)
    out[...] = bits(x)
    return out


def test_metadata_kernel_source(host):
    rng = np.random.default_rng(0)
    for sm, hpk in ((148, 16), (148, 128), (132, 64), (8, 16)):
        for b in (1, 3, 37, 128, 300):
            seqlens = rng.integers(1, 20000, size=b).astype(np.int32)
            parts = omla.num_sm_parts_for(hpk, 1, sm)
            meta, ns = aligned((parts, 8), np.int32), np.zeros(b + 1, np.int32)
            meta[...] = -1
            assert host.host_mla_get_metadata(_p(seqlens), b, parts, _p(meta), _p(ns)) == 0
            want, want_ns = omla.get_mla_metadata(seqlens, hpk, 1, sm)
            assert np.array_equal(meta, want) and np.array_equal(ns, want_ns)      # incl. the zeroed columns 5-7
            if ref_cpu.available():
                r_meta, r_ns = ref_cpu.get_mla_metadata(seqlens, hpk, 1, sm)
                assert np.array_equal(meta[:, :5], r_meta[:, :5]) and np.array_equal(ns, r_ns)


def test_quantiser_kernel_sources(host):
    rng = np.random.default_rng(1)
    q = bf16_rows(rng, (61, 3, 576))
    q[2] = 0
    rows = 61 * 3
    nope, scale, rope = aligned((61, 3, 512), np.uint8), np.zeros((61, 3, 1), np.float32), aligned((61, 3, 64), np.uint16)
    assert host.host_quantize_q(_p(q), rows, _p(nope), _p(scale), _p(rope)) == 0
    oc, osc, orr = omla.quantize_ckv_per_token_head(vals(q))
    assert np.array_equal(nope, oc) and np.array_equal(scale.view(np.uint32), osc.view(np.uint32)) and np.array_equal(rope, bits(orr))
    if ref_cpu.available():
        rc, rs, rr = ref_cpu.quantize_ckv_per_token_head(q)
        assert np.array_equal(nope, rc) and np.array_equal(scale.view(np.uint32), rs.view(np.uint32)) and np.array_equal(rope, rr)
    # quantise-and-cache from strided rows (the kv columns of a wider projection output), then indexed de-quantisation
    n, slots = 45, 64 * 5
    wide = bf16_rows(rng, (n, 2112))
    loc = rng.permutation(slots)[:n].astype(np.int32)
    kl, ks, kr = ref_cpu.pool_arrays(slots)
    assert host.host_quantize_and_cache_k(_p(wide[:, 1536:]), 2112, n, _p(loc), _p(kl), _p(ks), _p(kr)) == 0
    key = np.ascontiguousarray(wide[:, 1536:])
    ol, os_, orr = np.zeros((slots, 1, 512), np.uint8), np.zeros((slots, 1, 1), np.float32), np.zeros((slots, 1, 64), np.float32)
    omla.quantize_and_cache_k(vals(key)[:, None, :], ol, os_, orr, loc)
    assert np.array_equal(kl, ol) and np.array_equal(ks.view(np.uint32), os_.view(np.uint32)) and np.array_equal(kr, bits(orr))
    idx = rng.choice(loc, size=31).astype(np.int32)
    lo, ro = aligned((31, 1, 512), np.uint16), aligned((31, 1, 64), np.uint16)
    for cap in (-1, 5):   # free grid / grid fixed by max_num_tokens (CUDA-graph replay)
        lo[...] = 0
        assert host.host_dequantize_indexed(_p(kl), _p(kr), _p(ks), _p(idx), 31, cap, _p(lo), _p(ro)) == 0
        dl, dr = omla.dequantize_ckv_fused_indexed(ol, orr, os_, idx)
        assert np.array_equal(lo, bits(dl)) and np.array_equal(ro, bits(dr))
    if ref_cpu.available():
        rl, rs2, rr2 = ref_cpu.pool_arrays(slots)
        ref_cpu.quantize_and_cache_k(key[:, None, :], rl, rs2, rr2, loc)
        assert np.array_equal(kl, rl) and np.array_equal(ks.view(np.uint32), rs2.view(np.uint32)) and np.array_equal(kr, rr2)
        r_lo, r_ro = ref_cpu.dequantize_ckv_fused_indexed(rl, rr2, rs2, idx)
        assert np.array_equal(lo, r_lo) and np.array_equal(ro, r_ro)


def test_activation_quantiser_kernel_source(host):
    rng = np.random.default_rng(2)
    for m, k in ((1, 128), (5, 512), (37, 1536), (128, 256)):
        wide = bf16_rows(rng, (m, k + 576))
        wide[0, :128] = 0
        stride = (m + 3) // 4 * 4
        q, sc = aligned((m, k), np.uint8), np.zeros((k // 128, stride), np.float32)
        # a column slice of the wider rows, read in place (row stride k + 576, group stride 128)
        assert host.host_act_quant_1x128_strided(_p(wide), k + 576, 128, m, k, _p(q), _p(sc), stride, 1e-10) == 0
        x = np.ascontiguousarray(wide[:, :k])
        oq, os_ = ogemm.per_token_group_quant_fp8(vals(x))
        assert np.array_equal(q, oq) and np.array_equal(sc[:, :m].T.view(np.uint32), os_.view(np.uint32))
        if ref_cpu.available():
            rq, rs = ref_cpu.per_token_group_quant_fp8(x)
            assert np.array_equal(q, rq) and np.array_equal(np.ascontiguousarray(sc[:, :m].T).view(np.uint32), rs.view(np.uint32))
    # a per-head slice [m, heads, 128] of [m, heads, 512] (group stride 512)
    m, heads = 6, 16
    o = bf16_rows(rng, (m, heads, 512))
    stride = (m + 3) // 4 * 4
    q, sc = aligned((m, heads * 128), np.uint8), np.zeros((heads, stride), np.float32)
    assert host.host_act_quant_1x128_strided(_p(o), heads * 512, 512, m, heads * 128, _p(q), _p(sc), stride, 1e-10) == 0
    oq, os_ = ogemm.per_token_group_quant_fp8(vals(np.ascontiguousarray(o[:, :, :128]).reshape(m, heads * 128)))
    assert np.array_equal(q, oq) and np.array_equal(sc[:, :m].T.view(np.uint32), os_.view(np.uint32))


def test_combine_kernel_source(host):
    rng = np.random.default_rng(3)
    b, s_q, h, parts = 3, 2, 4, 8
    M = s_q * h
    ns = np.array([0, 2, 3, 8], np.int32)
    oacc = rng.standard_normal((b + parts, M, 512)).astype(np.float32)
    lacc = (rng.standard_normal((b + parts, M)) * 3 + 5).astype(np.float32)
    lacc[3, 1] = -np.inf
    lacc[4:8, 2] = -np.inf
    out, lse = aligned((b, M, 512), np.uint16), np.zeros((b, h, s_q), np.float32)
    assert host.host_mla_combine(_p(oacc), _p(lacc), _p(ns), _p(out), _p(lse), b, s_q, h) == 0
    want, want_lse = np.zeros((b, M, 512), np.float32), np.zeros((b, M), np.float32)
    omla.combine_partials(oacc, lacc, ns, want, want_lse)
    got_lse = lse.transpose(0, 2, 1).reshape(b, M)                      # [b, h, s_q] -> rows ordered (s, h)
    assert np.allclose(got_lse, want_lse, rtol=3e-7, atol=0)
    ulp = np.abs(out.astype(np.int32) - bits(want).astype(np.int32))
    assert ulp.max() <= 1 and (ulp != 0).mean() < 5e-3 and not out[1].any()
    if ref_cpu.available():
        r_out, r_lse = ref_cpu.mla_combine(oacc, lacc, ns, b, s_q, h, parts)
        assert np.allclose(got_lse, r_lse, rtol=3e-7, atol=0)
        ulp = np.abs(out.astype(np.int32) - r_out.astype(np.int32))
        assert ulp.max() <= 1 and (ulp != 0).mean() < 5e-3


@pytest.mark.parametrize("neox", [False, True])
def test_fused_prologue_kernel_source(host, neox):
    rng = np.random.default_rng(4 + int(neox))
    tokens, heads, slots, max_pos = 9, 5, 64 * 3, 512
    q_buf = bf16_rows(rng, (tokens, heads, 576))       # q_nope = [..., :512] of the absorb buffer
    q_full = bf16_rows(rng, (tokens, heads, 192))      # q_pe = [..., 128:] of q_b_proj's output
    qkv_a = bf16_rows(rng, (tokens, 2112))             # kv = columns 1536.. of the fused projection output
    q_buf[0, 0, :512] = 0
    inv_freq = 1.0 / (10000.0 ** (np.arange(0, 64, 2, dtype=np.float32) / 64))
    freqs = np.arange(max_pos, dtype=np.float32)[:, None] * inv_freq[None, :]
    cache = aligned((max_pos, 64), np.float32)
    cache[...] = np.concatenate([np.cos(freqs), np.sin(freqs)], -1).astype(np.float32)
    positions = rng.integers(0, max_pos, size=tokens).astype(np.int64)
    loc = rng.permutation(slots)[:tokens].astype(np.int32)
    kl, ks, kr = ref_cpu.pool_arrays(slots)
    nope, scale, rope = aligned((tokens, heads, 512), np.uint8), np.zeros((tokens, heads, 1), np.float32), aligned((tokens, heads, 64), np.uint16)
    q_pe, kv = q_full[:, :, 128:], qkv_a[:, 1536:]
    assert host.host_rope_quantize_fused(
        _p(q_buf), heads * 576, 576, ctypes.c_void_p(q_full.ctypes.data + 128 * 2), heads * 192, 192,
        ctypes.c_void_p(qkv_a.ctypes.data + 1536 * 2), 2112, _p(positions), _p(cache), int(neox), tokens, heads,
        _p(nope), _p(scale), _p(rope), _p(loc), _p(kl), _p(ks), _p(kr)) == 0
    ol, os_, orr = np.zeros((slots, 1, 512), np.uint8), np.zeros((slots, 1, 1), np.float32), np.zeros((slots, 1, 64), np.float32)
    wn, ws, wr = orope.rope_quantize_fused(vals(np.ascontiguousarray(q_buf[:, :, :512])), vals(np.ascontiguousarray(q_pe)),
                                           vals(np.ascontiguousarray(kv)), positions, cache, neox, loc, ol, os_, orr)
    assert np.array_equal(nope, wn) and np.array_equal(scale.view(np.uint32), ws.view(np.uint32)) and np.array_equal(rope, bits(wr))
    assert np.array_equal(kl, ol) and np.array_equal(ks.view(np.uint32), os_.view(np.uint32)) and np.array_equal(kr, bits(orr))
