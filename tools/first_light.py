"""GPU first-light diagnostics: UMMA descriptor lab, small kernels, GEMM and MLA decode vs the oracle.
Usage: python tools/first_light.py [stage ...]   stages: lab small gemm mla"""
import os, sys, time, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ctypes
import numpy as np
import torch
from sglang_fluentllm_b200 import _lib, flash_mla_fp8 as fm, deep_gemm as dg
from oracle import mla as omla, fp8 as ofp8, gemm as ogemm

dev = torch.device("cuda:0")
stages = sys.argv[1:] or ["lab", "small", "gemm", "mla"]


def lab():
    L = _lib.lib()
    for variant in (0, 1, 2, 3, 4):
        for n in (16, 64):
            err = ctypes.c_float(-1)
            rc = L.b200_umma_selftest(variant, n, ctypes.byref(err))
            print("lab variant %d n=%d rc=%d rel_err=%g %s" % (variant, n, rc, err.value,
                  L.b200_last_error().decode() if rc else ""), flush=True)


def small():
    torch.manual_seed(0)
    seqlens = torch.tensor([4096] * 128, dtype=torch.int32)
    for sl, hpk in ((seqlens, 16), (seqlens, 128), (torch.randint(1, 9000, (37,), dtype=torch.int32), 64)):
        meta, ns = fm.get_mla_metadata(sl.to(dev), hpk, 1)
        om, ons = omla.get_mla_metadata(sl.numpy(), hpk, 1, _lib.lib().b200_sm_count())
        print("metadata hpk=%d parts=%d meta_equal=%s splits_equal=%s" % (
            hpk, meta.shape[0], np.array_equal(meta.cpu().numpy()[:, :5], om[:, :5]), np.array_equal(ns.cpu().numpy(), ons)))
    q = torch.randn(8, 1, 16, 576, dtype=torch.bfloat16, device=dev) * 3
    nope, scale, rope = fm.quantize_ckv_per_token_head(q, 512)
    oc, osc, orp = omla.quantize_ckv_per_token_head(q.float().cpu().numpy(), 512)
    print("quantize_q codes_equal=%s scale_equal=%s rope_equal=%s" % (
        np.array_equal(nope.cpu().numpy(), oc), np.array_equal(scale.cpu().numpy(), osc),
        np.array_equal(rope.float().cpu().numpy(), orp)))
    x = torch.randn(33, 512, dtype=torch.bfloat16, device=dev)
    from sglang_fluentllm_b200 import fp8_kernel as fk
    xq, xs = fk.sglang_per_token_group_quant_fp8(x, 128, column_major_scales=True, scale_tma_aligned=True)
    oq, os_ = ogemm.per_token_group_quant_fp8(x.float().cpu().numpy())
    print("act_quant codes_equal=%s scales_equal=%s" % (
        np.array_equal(xq.view(torch.uint8).cpu().numpy(), oq), np.array_equal(xs.cpu().numpy(), os_)))


def gemm():
    torch.manual_seed(0)
    for (m, n, k) in ((16, 128, 128), (32, 256, 512), (128, 7168, 2048), (128, 4096, 7168), (100, 192, 256), (300, 256, 384),
                      (128, 2112, 7168), (64, 7168, 16384), (128, 24576, 1536), (7, 2112, 7168)):
        x = torch.randn(m, k, dtype=torch.bfloat16, device=dev)
        w = torch.randn(n, k, dtype=torch.bfloat16, device=dev)
        a8, a_s = dg.per_token_cast_to_fp8(x)
        w8, w_s = dg.per_block_cast_to_fp8(w)
        c = torch.zeros(m, n, dtype=torch.bfloat16, device=dev)
        dg.gemm_fp8_fp8_bf16_nt((a8, dg.get_col_major_tma_aligned_tensor(a_s)), (w8, w_s), c, False)
        torch.cuda.synchronize()
        ref = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(a8.view(torch.uint8).cpu().numpy()), a_s.cpu().numpy(),
                                          ofp8.e4m3_to_f32(w8.view(torch.uint8).cpu().numpy()), w_s.cpu().numpy())
        got = c.float().cpu().numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print("gemm m=%d n=%d k=%d rel=%.3e" % (m, n, k, rel), flush=True)


def make_case(b, h, s_q, seqlens, shuffle=True, nan_pad=True, seed=0):
    torch.manual_seed(seed)
    d, dv = 576, 512
    seqlens = torch.as_tensor(seqlens, dtype=torch.int32)
    max_blocks = int((int(seqlens.max()) + 63) // 64) + 1
    n_pages = b * max_blocks + 1
    q = torch.randn(b, s_q, h, d, dtype=torch.bfloat16, device=dev)
    kv = torch.randn(n_pages * 64, 1, d, dtype=torch.bfloat16, device=dev)
    perm = torch.randperm(n_pages - 1) + 1 if shuffle else torch.arange(1, n_pages)
    block_table = perm[: b * max_blocks].reshape(b, max_blocks).to(torch.int32)
    if nan_pad:
        kvv = kv.view(n_pages, 64, d)
        for i in range(b):
            sl = int(seqlens[i])
            blk, off = sl // 64, sl % 64
            for j in range(blk, max_blocks):
                pg = int(block_table[i, j])
                kvv[pg, (off if j == blk else 0):] = float("nan")
    k_lora = torch.zeros(n_pages * 64, 1, dv, dtype=torch.uint8, device=dev)
    k_scale = torch.zeros(n_pages * 64, 1, 1, dtype=torch.float32, device=dev)
    k_rope = torch.zeros(n_pages * 64, 1, d - dv, dtype=torch.bfloat16, device=dev)
    idx = torch.arange(n_pages * 64, dtype=torch.int32, device=dev)
    fm.quantize_and_cache_k(kv, k_lora, k_scale, k_rope, idx, dv)
    q_nope, q_scale, q_rope = fm.quantize_ckv_per_token_head(q, dv)
    return dict(q_nope=q_nope, q_rope=q_rope, q_scale=q_scale, k_lora=k_lora.view(n_pages, 64, 1, dv),
                k_rope=k_rope.view(n_pages, 64, 1, d - dv), k_scale=k_scale.view(n_pages, 64, 1, 1),
                block_table=block_table, seqlens=seqlens, n_pages=n_pages)


def run_case(name, b, h, s_q, seqlens, causal=False, faithful=True):
    c = make_case(b, h, s_q, seqlens)
    sl = c["seqlens"].to(dev)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1)
    t0 = time.time()
    out, lse = fm.flash_mla_ckv_fp8_per_token(c["q_nope"], c["q_rope"], c["k_lora"], c["k_rope"], c["q_scale"], c["k_scale"],
                                              c["block_table"].to(dev), sl, 512, meta, ns, causal=causal)
    torch.cuda.synchronize()
    args = (c["q_nope"].cpu().numpy(), c["q_rope"].float().cpu().numpy(), c["k_lora"].cpu().numpy(),
            c["k_rope"].float().cpu().numpy(), c["q_scale"].cpu().numpy(), c["k_scale"].cpu().numpy(),
            c["block_table"].numpy(), c["seqlens"].numpy())
    so, sl_ = omla.flash_mla_ckv_fp8_semantic(*args, causal=causal)
    got, gl = out.float().cpu().numpy(), lse.cpu().numpy()
    rel_s = np.linalg.norm(got - so) / max(np.linalg.norm(so), 1e-30)
    fin = np.isfinite(sl_)
    lse_err = np.max(np.abs(gl[fin] - sl_[fin])) if fin.any() else 0.0
    msg = "mla %-28s b=%d h=%d s_q=%d rel_semantic=%.3e lse_maxabs=%.3e nan_out=%d" % (
        name, b, h, s_q, rel_s, lse_err, int(np.isnan(got).sum()))
    if faithful:
        fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta.cpu().numpy(), ns.cpu().numpy(), causal=causal)
        msg += " rel_faithful=%.3e" % (np.linalg.norm(got - fo) / max(np.linalg.norm(fo), 1e-30))
    print(msg, flush=True)


def mla():
    run_case("single-short", 1, 16, 1, [100])
    run_case("single-2tiles", 1, 16, 1, [200])
    run_case("ragged-h16", 4, 16, 1, [300, 64, 1, 517])
    run_case("ragged-h32", 3, 32, 1, [129, 700, 65])
    run_case("ragged-h64", 3, 64, 1, [129, 700, 65])
    run_case("ragged-h128", 2, 128, 1, [333, 1025])
    run_case("long-split-h16", 2, 16, 1, [20000, 9000], faithful=False)
    run_case("sq2-causal-h16", 3, 16, 2, [300, 130, 64], causal=True)
    run_case("sq4-causal-h16", 2, 16, 4, [300, 65], causal=True)
    run_case("b128-s4096-h16", 128, 16, 1, [4096] * 128, faithful=False)


def mla_pair():
    """Shapes with M = h * s_q > 64: the CTA-pair kernel."""
    run_case("pair-single-short", 1, 128, 1, [100])
    run_case("pair-1page", 1, 128, 1, [64])
    run_case("pair-ragged-h128", 3, 128, 1, [333, 1025, 1])
    run_case("pair-h96", 2, 96, 1, [200, 77])
    run_case("pair-h64-sq2-causal", 3, 64, 2, [300, 130, 64], causal=True)
    run_case("pair-h128-sq2-causal", 2, 128, 2, [300, 129], causal=True)
    run_case("pair-h32-sq4-causal", 2, 32, 4, [300, 65], causal=True)
    run_case("pair-long-split-h128", 2, 128, 1, [20000, 9000], faithful=False)
    run_case("pair-b64-s2048-h128", 64, 128, 1, [2048] * 64, faithful=False)


for st in stages:
    try:
        if st in ("rate", "pair"):
            continue
        {"lab": lab, "small": small, "gemm": gemm, "mla": mla, "mla_pair": mla_pair}[st]()
    except Exception:
        traceback.print_exc()
        print("STAGE %s FAILED" % st, flush=True)


def rate():
    L = _lib.lib()
    names = {8: "fp8 K/K divergent issue", 0: "fp8 K/K", 1: "bf16 K/K", 2: "fp8 A-MN/B-MN(none)", 3: "fp8 A-MN/B-K", 4: "fp8 A-K/B-MN(none)"}
    for variant in (0, 1, 2, 3, 4, 8):
        for n in (16, 32, 64, 128, 256):
            c = ctypes.c_float(-1)
            rc = L.b200_umma_rate(variant, n, 2048, ctypes.byref(c))
            print("rate %-22s n=%3d: %7.1f cycles/MMA %s" % (names[variant], n, c.value, "" if rc == 0 else L.b200_last_error().decode()), flush=True)


if "rate" in stages:
    rate()


def pair():
    """CTA-pair (cta_group::2) lab: derive the TMEM layout of D from a raw dump."""
    import numpy as np
    L = _lib.lib()
    lut = np.array([_e4m3(i) for i in range(256)], dtype=np.float32)
    rng = np.random.default_rng(5)

    def rnd(shape):
        c = rng.integers(0, 256, size=shape, dtype=np.uint8)
        c[(c & 0x7F) >= 0x58] &= 0x47  # |v| < 16, no NaN
        return c

    for b_mode, n in ((0, 128), (0, 256), (0, 64), (1, 256)):
        a = rnd((128, 128))
        b = rnd((n, 128)) if b_mode == 0 else rnd((128, n))
        out = np.zeros((2, 128, 256), dtype=np.float32)
        rc = L.b200_umma_pair_lab(b_mode, n, a.ctypes.data, b.ctypes.data, out.ctypes.data)
        if rc != 0:
            print("pair b_mode=%d n=%d FAILED: %s" % (b_mode, n, L.b200_last_error().decode()), flush=True)
            continue
        af, bf = lut[a], lut[b]
        d = af @ bf.T if b_mode == 0 else af @ bf  # [128][n]
        print("pair b_mode=%d n=%d" % (b_mode, n), flush=True)
        for cta in range(2):
            for l0 in range(0, 128, 32):
                desc = []
                for lane in (l0, l0 + 31):
                    x = out[cta, lane]
                    written = np.nonzero(x != -12345.0)[0]
                    if len(written) == 0:
                        desc.append("lane %d: untouched" % lane)
                        continue
                    c_lo, c_hi = written.min(), written.max() + 1
                    seg = x[c_lo:c_hi]
                    best = None
                    for row in range(128):
                        for off in range(0, n - len(seg) + 1, 8):
                            err = np.abs(d[row, off:off + len(seg)] - seg).max()
                            if best is None or err < best[0]:
                                best = (err, row, off)
                    desc.append("lane %d: tmem cols [%d,%d) = D[row %d, cols %d..%d) err %.2g" % (
                        lane, c_lo, c_hi, best[1], best[2], best[2] + len(seg), best[0]))
                print("  cta %d  %s" % (cta, " | ".join(desc)), flush=True)


def _e4m3(c):
    import math
    s, e, m = c >> 7, (c >> 3) & 0xF, c & 7
    if e == 0xF and m == 7:
        return float("nan")
    v = math.ldexp(m / 8.0, -6) if e == 0 else math.ldexp(1 + m / 8.0, e - 7)
    return -v if s else v


if "pair" in stages:
    pair()
