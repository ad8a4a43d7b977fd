"""CPU simulation of the kernels' P arithmetic (reference page codes, page weight carried by an e4m3 hi + lo pair) against
the faithful oracle.  Development aid: run `python tools/sim_hilo.py`.  Not used by the product."""
import math
import sys
import os

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import mla as omla
from oracle.fp8 import e4m3_to_f32, f32_to_e4m3, round_to_bf16


def hilo_attention(qn, qr, qs, kn_pages, kr_pages, ks_pages, block_ids, seqlen, border, scale_log2, headroom=4.0, use_lo=True,
                   half=True):
    """One request, all rows; sequential pages; lazy reference exponent R per row."""
    M = qn.shape[0]
    R = np.full(M, -np.inf, np.float32)
    O = np.zeros((M, 512), np.float32)
    L = np.zeros(M, np.float32)
    f = (qs * scale_log2).astype(np.float32)
    tok0 = 0
    LOG448 = np.float32(math.log2(448.0))
    for pid in block_ids:
        kn = e4m3_to_f32(kn_pages[pid]); kr = kr_pages[pid]; ks = ks_pages[pid].astype(np.float32)
        tok = tok0 + np.arange(64)
        vrow = tok < seqlen
        knv = np.where(vrow[:, None], kn, 0).astype(np.float32)
        krv = np.where(vrow[:, None], kr, 0).astype(np.float32)
        ksv = np.where(vrow, ks, 1).astype(np.float32)
        s = (qn @ knv.T + qr @ krv.T).astype(np.float32)
        ok = (tok[None, :] < border[:, None]) & vrow[None, :]
        y = (s * (f[:, None] * ksv[None, :]) + np.log2(ksv)[None, :]).astype(np.float32)
        y = np.where(ok, y, -np.inf)
        ymax = y.max(axis=1)
        live = np.isfinite(ymax)
        ym = np.where(live, ymax, 0).astype(np.float32)
        v = np.where(ok, np.exp2((y - ym[:, None] + LOG448).astype(np.float32)), 0).astype(np.float32)
        code = e4m3_to_f32(f32_to_e4m3(v))
        esum = (v / ksv[None, :]).sum(axis=1, dtype=np.float32)
        # reference move
        need = live & (ymax > R)
        Rn = np.where(need, ymax + np.float32(headroom), R).astype(np.float32)
        fac = np.where(np.isfinite(R), np.exp2(R - Rn), 0).astype(np.float32)
        fac = np.where(need, fac, 1).astype(np.float32)
        O *= fac[:, None]; L *= fac
        R = Rn
        w = np.where(live, np.exp2((ym - R).astype(np.float32)), 0).astype(np.float32)
        if half:
            wh = w.astype(np.float16)
            ph = (code.astype(np.float16) * wh[:, None]).astype(np.float16)
            w32 = wh.astype(np.float32)
            pf = ph.astype(np.float32)
        else:
            w32 = w
            pf = (code * w[:, None]).astype(np.float32)
        hi = e4m3_to_f32(f32_to_e4m3(pf))
        lo = e4m3_to_f32(f32_to_e4m3((pf - hi).astype(np.float32))) if use_lo else 0
        O += (hi @ knv + (lo @ knv if use_lo else 0)).astype(np.float32)
        L += w32 * esum
        tok0 += 64
    bad = (L == 0) | np.isnan(L)
    out = O / np.where(bad, 1, L)[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        lse = (np.log2(L) + R - LOG448) / np.float32(math.log2(math.e))
    return round_to_bf16(out.astype(np.float32)), np.where(bad, np.inf, lse)


def main():
    rng = np.random.default_rng(0)
    b, h, s_q = 3, 16, 1
    seqlens = np.array([300, 1000, 65], np.int32)
    maxb = 17
    n_pages = b * maxb + 1
    q = round_to_bf16(rng.standard_normal((b, s_q, h, 576)).astype(np.float32))
    kv = round_to_bf16(rng.standard_normal((n_pages * 64, 1, 576)).astype(np.float32))
    qn, qs, qr = omla.quantize_ckv_per_token_head(q)
    kl = np.zeros((n_pages * 64, 1, 512), np.uint8); ks = np.zeros((n_pages * 64, 1, 1), np.float32); kr = np.zeros((n_pages * 64, 1, 64), np.float32)
    omla.quantize_and_cache_k(kv, kl, ks, kr, np.arange(n_pages * 64))
    kl = kl.reshape(n_pages, 64, 1, 512); ks = ks.reshape(n_pages, 64, 1, 1); kr = kr.reshape(n_pages, 64, 1, 64)
    bt = (rng.permutation(n_pages - 1)[: b * maxb] + 1).reshape(b, maxb).astype(np.int32)
    meta, ns = omla.get_mla_metadata(seqlens, s_q * h, 1, 148)
    # single-part metadata so the oracle sees whole requests (combine path not under test here)
    meta1 = np.zeros((1, 8), np.int32); meta1[0, :5] = [0, 0, b - 1, seqlens[-1], 0]
    ns1 = np.arange(b + 1, dtype=np.int32)
    fo, fl = omla.flash_mla_ckv_fp8_per_token(qn, qr, kl, kr, qs, ks, bt, seqlens, meta1, ns1)
    so, sl = omla.flash_mla_ckv_fp8_semantic(qn, qr, kl, kr, qs, ks, bt, seqlens)
    scale_log2 = np.float32(np.float32(576 ** -0.5) * np.float32(math.log2(math.e)))
    rel = lambda a, r: float(np.linalg.norm(a - r) / np.linalg.norm(r))
    for name, kw in [("hi+lo fp16 w", {}), ("hi+lo f32 w", dict(half=False)), ("hi only", dict(use_lo=False))]:
        out = np.zeros_like(fo); lse = np.zeros_like(fl)
        for i in range(b):
            nb = (int(seqlens[i]) + 63) // 64
            M = s_q * h
            o, l = hilo_attention(e4m3_to_f32(qn[i].reshape(M, 512)), qr[i].reshape(M, 64), qs[i].reshape(M),
                                  kl[:, :, 0, :], kr[:, :, 0, :], ks[:, :, 0, 0], bt[i, :nb], int(seqlens[i]),
                                  np.full(M, int(seqlens[i])), scale_log2, **kw)
            out[i] = o.reshape(s_q, h, 512); lse[i] = l.reshape(s_q, h).T
        print("%-14s rel vs faithful %.3e   vs semantic %.3e   lse max abs diff %.2e" % (name, rel(out, fo), rel(out, so), np.abs(lse - fl).max()))
    print("faithful vs semantic: %.3e" % rel(fo, so))


if __name__ == "__main__":
    main()
