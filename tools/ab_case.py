"""A/B of two builds of the MLA decode kernels on one box: time at the metric shape and distance from the
reference-faithful oracle on the same seeded inputs.  Usage: python tools/ab_case.py <root of the tree to load the package
from>  (the oracle is always this repo's).  Prints one JSON line."""
import json
import os
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
root = os.path.abspath(sys.argv[1]) if len(sys.argv) > 1 else HERE
sys.path.insert(0, root)
import numpy as np
import torch
from sglang_fluentllm_b200 import flash_mla_fp8 as fm, _lib

sys.path.insert(0, HERE)
if root != HERE:  # load this repo's oracle under its own name even when another tree's package is first on the path
    import importlib.util
    spec = importlib.util.spec_from_file_location("oracle", os.path.join(HERE, "oracle", "__init__.py"),
                                                  submodule_search_locations=[os.path.join(HERE, "oracle")])
    mod = importlib.util.module_from_spec(spec)
    sys.modules["oracle"] = mod
    spec.loader.exec_module(mod)
from oracle import mla as omla

# the tile schedule the backend runs (per-request cost tuned for B200); a tree without the extension keeps the reference's
SCHED = lambda m: ({"fixed_overhead_blocks": fm.tuned_overhead_blocks(m), **({"num_sm_parts": fm.tuned_num_sm_parts(m)} if hasattr(fm, "tuned_num_sm_parts") and not os.environ.get("B200_REF_PARTS") else {})} if hasattr(fm, "tuned_overhead_blocks") else {})

dev = torch.device("cuda:0")


def case(b, h, s_q, seqlens, seed=0):
    torch.manual_seed(seed)
    seqlens = torch.as_tensor(seqlens, dtype=torch.int32)
    max_blocks = int((int(seqlens.max()) + 63) // 64) + 1
    n_pages = b * max_blocks + 1
    q = torch.randn(b, s_q, h, 576, dtype=torch.bfloat16, device=dev)
    kv = torch.randn(n_pages * 64, 1, 576, dtype=torch.bfloat16, device=dev)
    bt = (torch.randperm(n_pages - 1)[: b * max_blocks] + 1).reshape(b, max_blocks).to(torch.int32)
    kl = torch.zeros(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev)
    ks = torch.zeros(n_pages * 64, 1, 1, dtype=torch.float32, device=dev)
    kr = torch.zeros(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
    fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=dev), 512)
    qn, qs, qr = fm.quantize_ckv_per_token_head(q, 512)
    return qn, qr, qs, kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), ks.view(n_pages, 64, 1, 1), bt, seqlens


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def accuracy(b, h, s_q, seqlens):
    qn, qr, qs, kl, kr, ks, bt, sl = case(b, h, s_q, seqlens)
    sld = sl.to(dev)
    meta, ns = fm.get_mla_metadata(sld, s_q * h, 1, **SCHED(s_q * h))
    out, lse = fm.flash_mla_ckv_fp8_per_token(qn, qr, kl, kr, qs, ks, bt.to(dev), sld, 512, meta, ns, causal=s_q > 1)
    torch.cuda.synchronize()
    args = (qn.cpu().numpy(), qr.float().cpu().numpy(), kl.cpu().numpy(), kr.float().cpu().numpy(), qs.cpu().numpy(),
            ks.cpu().numpy(), bt.numpy(), sl.numpy())
    fo, fl = omla.flash_mla_ckv_fp8_per_token(*args, meta.cpu().numpy(), ns.cpu().numpy(), causal=s_q > 1)
    so, _ = omla.flash_mla_ckv_fp8_semantic(*args, causal=s_q > 1)
    o = out.float().cpu().numpy()
    return {"vs_faithful": round(rel(o, fo), 5), "vs_fp64_attention": round(rel(o, so), 5),
            "faithful_vs_fp64_attention": round(rel(fo, so), 5), "lse_max_abs": round(float(np.abs(lse.cpu().numpy() - fl).max()), 6)}


def timing(b, seq, h, s_q=1):
    pages_per = (seq + 63) // 64 + 1
    n_pages = b * pages_per + 1
    n_rot = max(2, min(8, int(2.5e8 // (n_pages * 64 * 644)) + 2))
    pools = []
    tmp = torch.empty(n_pages * 64, 1, 576, dtype=torch.bfloat16, device=dev)
    idx = torch.arange(n_pages * 64, dtype=torch.int32, device=dev)
    for _ in range(n_rot):
        tmp.normal_()
        kl = torch.empty(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev)
        ks = torch.empty(n_pages * 64, 1, 1, dtype=torch.float32, device=dev)
        kr = torch.empty(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
        fm.quantize_and_cache_k(tmp, kl, ks, kr, idx, 512)
        pools.append((kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), ks.view(n_pages, 64, 1, 1)))
    q = torch.randn(b, s_q, h, 576, dtype=torch.bfloat16, device=dev)
    qn, qs, qr = fm.quantize_ckv_per_token_head(q, 512)
    bt = (torch.randperm(n_pages - 1)[: b * pages_per].reshape(b, pages_per) + 1).to(torch.int32).to(dev)
    sl = torch.full((b,), seq, dtype=torch.int32, device=dev)
    meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, **SCHED(s_q * h))
    L = _lib.lib()
    out = torch.empty(b, s_q, h, 512, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(b, h, s_q, dtype=torch.float32, device=dev)
    ws = torch.empty(L.b200_mla_workspace_bytes(b, s_q, h, meta.shape[0]), dtype=torch.uint8, device=dev)
    s = torch.cuda.Stream()
    torch.cuda.synchronize()

    def fn(i):
        kl, kr, ks = pools[i % n_rot]
        fm.flash_mla_ckv_fp8_per_token(qn, qr, kl, kr, qs, ks, bt, sl, 512, meta, ns, causal=s_q > 1, out=out, lse=lse, workspace=ws)

    with torch.cuda.stream(s):
        for i in range(2):
            fn(i)
        s.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n_rot):
                fn(i)
        g.replay()
        s.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        for _ in range(5):
            g.replay()
        e1.record(s)
        s.synchronize()
    return round(e0.elapsed_time(e1) / (5 * n_rot) * 1e3, 1)


res = {"tree": root,
       "us": {"b128_seq4096_h128": timing(128, 4096, 128), "b128_seq4096_h16": timing(128, 4096, 16),
              "b64_seq4096_h16_sq4": timing(64, 4096, 16, 4), "b256_seq8192_h16": timing(256, 8192, 16)},
       "rel": {"h128_b4": accuracy(4, 128, 1, [1024, 700, 333, 65]), "h16_b4": accuracy(4, 16, 1, [1024, 700, 333, 65]),
               "h16_sq4_causal": accuracy(3, 16, 4, [600, 300, 130])}}
print(json.dumps(res))
