"""Pipeline timeline of the CTA-pair MLA kernel (part 0): run with B200_MLA_TRACE=65 (the default here).
Usage: python tools/trace_mla.py [H B SEQ]; also checks run-to-run determinism."""
import os, sys
os.environ.setdefault("B200_MLA_TRACE", "65")  # bit 0: timestamps, bit 6: the V-landing observer (big shapes only)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sglang_fluentllm_b200 import flash_mla_fp8 as fm, _lib
h = int(sys.argv[1]) if len(sys.argv) > 1 else 128
b = int(sys.argv[2]) if len(sys.argv) > 2 else 128
seq = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
dev = torch.device("cuda:0"); torch.manual_seed(0)
pp = seq // 64 + 1; n_pages = b * pp + 1
kv = torch.randn(n_pages * 64, 1, 576, dtype=torch.bfloat16, device=dev)
kl = torch.empty(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev); ks = torch.empty(n_pages * 64, 1, 1, device=dev); kr = torch.empty(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=dev), 512)
q = torch.randn(b, 1, h, 576, dtype=torch.bfloat16, device=dev)
qn, qs, qr = fm.quantize_ckv_per_token_head(q, 512)
bt = (torch.randperm(n_pages - 1)[: b * pp].reshape(b, pp) + 1).to(torch.int32).to(dev)
if os.environ.get("SHARED_PAGES"):  # diagnostic: every request walks the same few pages -> the KV stream is L2-resident
    n_sh = int(os.environ["SHARED_PAGES"])
    bt = ((torch.arange(b * pp).reshape(b, pp) % n_sh) + 1).to(torch.int32).to(dev)
sl = torch.full((b,), seq, dtype=torch.int32, device=dev)
meta, ns = fm.get_mla_metadata(sl, h, 1, fixed_overhead_blocks=fm.tuned_overhead_blocks(h))
outs = []
for _ in range(5):
    o, l = fm.flash_mla_ckv_fp8_per_token(qn, qr, kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), qs, ks.view(n_pages, 64, 1, 1), bt, sl, 512, meta, ns, softmax_scale=192 ** -0.5)
    torch.cuda.synchronize()
    outs.append((o.float().cpu().numpy(), l.cpu().numpy()))
for i in range(1, len(outs)):
    d = np.abs(outs[i][0] - outs[0][0])
    print("determinism run %d vs 0: out differs in %d elements (max %.3e), lse differs in %d" % (
        i, int((d > 0).sum()), float(d.max()), int((outs[i][1] != outs[0][1]).sum())))
    rows = np.argwhere(d.reshape(b, h, 512).max(axis=2) > 0)
    groups = {}
    for bb, hh in rows:
        groups.setdefault((int(bb), int(hh) // 64), []).append(int(hh))
    for (bb, half), hs in sorted(groups.items()):
        ld = np.abs(outs[i][1][bb, hs, 0] - outs[0][1][bb, hs, 0]).max()
        dm = d.reshape(b, h, 512)[bb, hs].max()
        if dm < 2e-3:
            continue
        blocks = "".join("X" if d.reshape(b, h, 16, 32)[bb, hs, k].max() > 2e-3 else "." for k in range(16))
        quads = sorted(set((x % 64) // 32 for x in hs))
        print("    request %3d rows %3d..%3d (%2d rows, CTA %d, local quadrant pairs %s): max |d out| %.3e  max |d lse| %.3e  d-blocks(32) %s" % (
            bb, min(hs), max(hs), len(hs), half, quads, float(dm), float(ld), blocks))
L = _lib.lib()
T = 64
buf = np.zeros(2 * 5 * T * 4, dtype=np.uint64)
_lib.check(L.b200_mla_read_trace(buf.ctypes.data, buf.size), "read_trace")
tr = buf.reshape(2, 5, T, 4).astype(np.int64)
for rank in range(2):
    t0 = tr[rank, 1, 0, 0] if rank == 0 else tr[rank, 0, 0, 0]
    print("rank %d (cycles relative to first event; d = delta to previous tile)" % rank)
    print(" tile | prod: Kissue Vwait  Vissue | mma: kwait  QKiss  pwait  PViss | softmax: swait  Srdy  computed arrived | period")
    prev = None
    for g in range(min(T, 40)):
        r = tr[rank, :, g, :] - t0
        if tr[rank, 2, g, 3] == 0:
            break
        ref = r[2, 3]
        per = "" if prev is None else "%6d" % (ref - prev)
        prev = ref
        print(" %4d | %7d %7d %7d | %7d %7d %7d %7d | %7d %7d %7d %7d | %s" % (
            g, r[0, 0], r[0, 1], r[0, 2], r[1, 0], r[1, 1], r[1, 2], r[1, 3], r[2, 0], r[2, 1], r[2, 2], r[2, 3], per))

    print(" softmax detail (deltas): tile | top->swait  swait->Srdy  Srdy->pass  pass->fin  fin->voted(computed)  computed->vland  vland->arrived")
    for g in range(8, 24):
        a = tr[rank, 2, g, :]; d = tr[rank, 4, g, :]
        print("   %3d | %6d %6d %6d %6d %6d %6d %6d" % (g, a[0] - d[3], a[1] - a[0], d[0] - a[1], d[1] - d[0], a[2] - d[1], d[2] - a[2], a[3] - d[2]))
    print(" K / MMA detail: tile | Kissue->landed  landed->QKstart  QKstart->QKissued  QKissued->Srdy | P arrived(this CTA)->PVstart  PVstart->PVissued | PVissued(g-4)->QKstart(g)  QKissued(g)->Kissue(g+2)")
    for g in range(8, 28):
        pr = tr[rank, 0, g, :]; m = tr[rank, 1, g, :]; a = tr[rank, 2, g, :]
        print("   %3d | %6d %6d %6d %6d | %6d %6d | %6d %6d" % (g, pr[3] - pr[0], m[0] - pr[3], m[1] - m[0], a[1] - m[1], m[2] - a[3], m[3] - m[2], m[0] - tr[rank, 1, g - 4, 3], tr[rank, 0, g + 2, 0] - m[1]))
    if rank == 1:
        print(" V detail (non-leader CTA): tile | Vwait(start)->Vissue   Vissue->landed   landed->P arrived(this CTA)")
        for g in range(8, 24):
            pr = tr[rank, 0, g, :]; a = tr[rank, 2, g, :]
            print("   %3d | %6d %6d %6d" % (g, pr[2] - pr[1], pr[3] - pr[2], a[3] - pr[3]))
    print(" request | epilogue: start  o_ready  stored | next request's first tile top (previous row's request)")
    for sgi in range(4):
        e = tr[rank, 3, sgi, :] - t0
        print(" %7d | %9d %8d %8d | %8d" % (sgi, e[0], e[1], e[2], e[3]))
    print(" request | epilogue detail (group 0, warp 0): first O chunk in registers  first chunk stored  last chunk read (O released)  last chunk stored | P.V issuer: first P of the request complete  O free")
    for sgi in range(4):
        e = tr[rank, 3, 8 + sgi, :] - t0
        m = tr[rank, 3, 16 + sgi, :] - t0
        print(" %7d | %9d %8d %8d %8d | %9d %8d" % (sgi, e[0], e[1], e[2], e[3], m[0], m[1]))

pt = np.zeros(2 * 128 * 4, dtype=np.uint64)
_lib.check(L.b200_mla_read_part_times(pt.ctypes.data, pt.size), "read_part_times")
pt = pt.reshape(128, 2, 4).astype(np.int64)
clk = (pt[:, 0, 3] - pt[:, 0, 2]) / np.maximum(pt[:, 0, 1] - pt[:, 0, 0], 1)  # SM cycles per ns, leader CTA of each part
print("SM clock during the kernel (clock64 / globaltimer, per part): median %.3f GHz, min %.3f, max %.3f; busy time median %.1f us" % (
    np.median(clk[:meta.shape[0] - 1]), clk[:meta.shape[0] - 1].min(), clk[:meta.shape[0] - 1].max(),
    np.median(pt[:meta.shape[0] - 1, 0, 1] - pt[:meta.shape[0] - 1, 0, 0]) / 1e3))
nparts = meta.shape[0]
t_first = pt[:nparts, :, 0].min()
print("per-part timeline (us relative to the first CTA's entry): part | entry r0 r1 | softmax end r0 r1 | busy")
for part in range(nparts):
    e = (pt[part, :, 0] - t_first) / 1e3; x = (pt[part, :, 1] - t_first) / 1e3
    print("  %3d | %6.1f %6.1f | %6.1f %6.1f | %6.1f   meta %s" % (part, e[0], e[1], x[0], x[1], max(x) - min(e), meta[part, :5].tolist()))
