"""GPU tests of the drop-in boundary above the C ABI: the AttentionBackend plug-in driven the
way the reference's model runner drives it (eager decode, CUDA-graph capture + replay, speculative TARGET_VERIFY), and
the FP8 linear method.  The checker is fp32 torch attention over the de-quantised pool -- what the reference's
torch-native backend computes (oracle/pool.py::native_sdpa_decode restates it; golden in tests/golden/)."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import fp8 as ofp8
from oracle import gemm as ogemm
from helpers import quantize_rhs_128x128

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SCALING = 192 ** -0.5  # DeepSeek-V2/V3 MLA: (qk_nope 128 + rope 64) ** -0.5


def _runner(heads, draft_tokens=0, max_bs=8, context_len=2048, quant_method="per_token_head"):
    from sglang_fluentllm_b200 import flashmla_backend as fb
    alloc = fb.KVAllocator(size=64 * 256, device=DEV, max_batch_size=max_bs, max_context_len=context_len)
    pool = fb.MLATokenToKVPool(size=64 * 256, model_dtype=torch.bfloat16, kv_lora_rank=512, qk_rope_head_dim=64,
                               layer_num=2, device=DEV, quant_method=quant_method)
    cfg = SimpleNamespace(scaling=SCALING, num_attention_heads=heads, kv_lora_rank=512, qk_nope_head_dim=128,
                          qk_rope_head_dim=64, v_head_dim=512, context_len=context_len)
    sa = SimpleNamespace(speculative_algorithm="EAGLE" if draft_tokens else None, speculative_num_draft_tokens=draft_tokens)
    fp8_kv = quant_method == "per_token_head"
    mr = SimpleNamespace(model_config=cfg, server_args=sa,
                         kv_cache_dtype=torch.float8_e4m3fn if fp8_kv else torch.bfloat16,
                         kv_cache_quant_method=quant_method, kv_allocator=alloc, dtype=torch.bfloat16, device=DEV)
    layer = SimpleNamespace(layer_id=1, tp_q_head_num=heads, head_dim=576, v_head_dim=512, scaling=SCALING)
    return fb, mr, alloc, pool, layer


def _fill_context(alloc, pool, layer, ctx_lens, seed=0):
    """Prefill stand-in: random latent K for every request's context, written through the pool's own set_kv_buffer."""
    torch.manual_seed(seed)
    locs = []
    for i, n in enumerate(ctx_lens):
        loc = alloc.alloc(i, n, 0)
        assert loc is not None
        pool.set_kv_buffer(layer, loc, torch.randn(n, 1, 576, dtype=torch.bfloat16, device=DEV))
        locs.append(loc)
    return locs


def _reference(pool, layer, q, all_locs, n_q, causal_tail):
    """fp32 attention over the de-quantised pool rows of each request; q [bs, n_q, h, 576].  Row j of the n_q query
    tokens sees all but the last (n_q - 1 - j) keys (speculative verify: causal over the draft tokens)."""
    outs = []
    for i, loc in enumerate(all_locs):
        lora, rope = pool.get_key_split_contiguous(layer.layer_id, loc)
        k = torch.cat([lora, rope], dim=-1).float().squeeze(1)     # [n, 576]
        v = lora.float().squeeze(1)                                 # [n, 512]
        o = []
        for j in range(n_q):
            n_vis = k.shape[0] - (n_q - 1 - j if causal_tail else 0)
            logits = (q[i, j].float() @ k[:n_vis].t()) * SCALING    # [h, n_vis]
            o.append(torch.softmax(logits, dim=-1) @ v[:n_vis])
        outs.append(torch.stack(o))
    return torch.stack(outs)                                        # [bs, n_q, h, 512]


def _rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


@pytest.mark.parametrize("heads", [16, 128])
def test_backend_forward_decode_eager(heads):
    fb, mr, alloc, pool, layer = _runner(heads)
    ctx = [130, 64, 700, 1]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    backend = fb.FlashMLABackend(mr)
    new_loc = torch.cat([alloc.alloc(i, 1, n) for i, n in enumerate(ctx)])
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=torch.tensor([n + 1 for n in ctx], device=DEV),
                             req_pool_indices=torch.arange(bs, device=DEV), out_cache_loc=new_loc,
                             token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DECODE)
    backend.init_forward_metadata(fbatch)
    torch.manual_seed(1)
    q = torch.randn(bs, heads * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs, 1, 576, dtype=torch.bfloat16, device=DEV)
    out = backend.forward(q, k_new, k_new[..., :512], layer, fbatch)   # base-class dispatch on forward_mode
    assert out.shape == (bs, heads * 512) and out.dtype == torch.bfloat16
    all_locs = [torch.cat([ctx_locs[i], new_loc[i:i + 1]]) for i in range(bs)]
    ref = _reference(pool, layer, q.view(bs, 1, heads, 576), all_locs, 1, False)
    assert _rel(out.view(bs, 1, heads, 512), ref) < 5e-2
    # the block table the kernel walked is the allocator's gather, bit for bit
    assert torch.equal(backend.forward_metadata.block_table, alloc.req_to_page[fbatch.req_pool_indices])


@pytest.mark.parametrize("heads,draft", [(16, 0), (128, 0), (16, 4)])
def test_backend_bf16_pool(heads, draft):
    """--kv-cache-quant-method none: bf16 pool, the backend's third branch (flashmla_backend.py:240-256) -> the bf16
    kernel.  Decode and speculative target-verify, against fp32 attention over the pool rows (bf16 tolerance)."""
    fb, mr, alloc, pool, layer = _runner(heads, draft_tokens=draft, quant_method="none")
    ctx = [130, 64, 700, 1]
    bs, n_q = len(ctx), max(draft, 1)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    backend = fb.FlashMLABackend(mr)
    new_loc = torch.cat([alloc.alloc(i, n_q, n) for i, n in enumerate(ctx)])
    mode = fb.ForwardMode.TARGET_VERIFY if draft else fb.ForwardMode.DECODE
    seq_lens = torch.tensor(ctx if draft else [n + 1 for n in ctx], device=DEV)
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=seq_lens, req_pool_indices=torch.arange(bs, device=DEV),
                             out_cache_loc=new_loc, token_to_kv_pool=pool, forward_mode=mode)
    backend.init_forward_metadata(fbatch)
    torch.manual_seed(1)
    q = torch.randn(bs * n_q, heads * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs * n_q, 1, 576, dtype=torch.bfloat16, device=DEV)
    out = backend.forward(q, k_new, k_new[..., :512], layer, fbatch)
    assert out.shape == (bs * n_q, heads * 512) and out.dtype == torch.bfloat16
    all_locs = [torch.cat([ctx_locs[i], new_loc[i * n_q:(i + 1) * n_q]]) for i in range(bs)]
    ref = _reference(pool, layer, q.view(bs, n_q, heads, 576), all_locs, n_q, bool(draft))
    assert _rel(out.view(bs, n_q, heads, 512), ref) < 4e-3


def test_backend_cuda_graph_capture_and_replay():
    """The runner's protocol (cuda_graph_runner.py:405-520): static state sized for max_bs, capture at one set of
    lengths, replay after the sequences grew with only the metadata refreshed outside the graph."""
    heads = 16
    fb, mr, alloc, pool, layer = _runner(heads)
    ctx = [300, 65, 129]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    backend = fb.FlashMLABackend(mr)
    backend.init_cuda_graph_state(max_bs=4)
    assert backend.get_cuda_graph_seq_len_fill_value() == 1
    req = torch.arange(bs, device=DEV)
    torch.manual_seed(2)
    q = torch.randn(bs, heads * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs, 1, 576, dtype=torch.bfloat16, device=DEV)
    seq_lens = torch.tensor([n + 1 for n in ctx], device=DEV)
    out_loc = torch.cat([alloc.alloc(i, 1, n) for i, n in enumerate(ctx)])
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=seq_lens, req_pool_indices=req, out_cache_loc=out_loc,
                             token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DECODE)
    backend.init_forward_metadata_capture_cuda_graph(bs, bs, req, seq_lens, fb.ForwardMode.DECODE, None)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(2):  # warm-up runs before capture, as the runner does
            backend.forward_decode(q, k_new, k_new[..., :512], layer, fbatch)
        s.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=s):
            out_static = backend.forward_decode(q, k_new, k_new[..., :512], layer, fbatch)
        graph.replay()
        s.synchronize()
    locs1 = [torch.cat([ctx_locs[i], out_loc[i:i + 1]]) for i in range(bs)]
    assert _rel(out_static.view(bs, 1, heads, 512), _reference(pool, layer, q.view(bs, 1, heads, 576), locs1, 1, False)) < 5e-2
    # next step: every sequence grew by one token; new q / k / slots go into the SAME static tensors
    out_loc2 = torch.cat([alloc.alloc(i, 1, n + 1) for i, n in enumerate(ctx)])
    q.copy_(torch.randn_like(q))
    k_new.copy_(torch.randn_like(k_new))
    fbatch.out_cache_loc.copy_(out_loc2)
    seq_lens.add_(1)
    backend.init_forward_metadata_replay_cuda_graph(bs, req, seq_lens, int(seq_lens.sum()), fb.ForwardMode.DECODE, None,
                                                    seq_lens_cpu=seq_lens.cpu())
    with torch.cuda.stream(s):
        graph.replay()
        s.synchronize()
    locs2 = [torch.cat([locs1[i], out_loc2[i:i + 1]]) for i in range(bs)]
    assert _rel(out_static.view(bs, 1, heads, 512), _reference(pool, layer, q.view(bs, 1, heads, 576), locs2, 1, False)) < 5e-2


def test_backend_target_verify_is_causal_over_draft_tokens():
    """Speculative verify (flashmla_backend.py:105-176): s_q = draft tokens, attention over seq_lens + draft tokens,
    causal among the drafts."""
    heads, draft = 16, 2
    fb, mr, alloc, pool, layer = _runner(heads, draft_tokens=draft)
    ctx = [200, 77]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    backend = fb.FlashMLABackend(mr)
    new_loc = torch.cat([alloc.alloc(i, draft, n) for i, n in enumerate(ctx)])
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=torch.tensor(ctx, device=DEV), req_pool_indices=torch.arange(bs, device=DEV),
                             out_cache_loc=new_loc, token_to_kv_pool=pool, forward_mode=fb.ForwardMode.TARGET_VERIFY)
    backend.init_forward_metadata(fbatch)
    torch.manual_seed(3)
    q = torch.randn(bs * draft, heads * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs * draft, 1, 576, dtype=torch.bfloat16, device=DEV)
    out = backend.forward(q, k_new, k_new[..., :512], layer, fbatch)
    assert out.shape == (bs * draft, heads * 512)
    all_locs = [torch.cat([ctx_locs[i], new_loc[i * draft:(i + 1) * draft]]) for i in range(bs)]
    ref = _reference(pool, layer, q.view(bs, draft, heads, 576), all_locs, draft, True)
    assert _rel(out.view(bs, draft, heads, 512), ref) < 5e-2
    with pytest.raises(NotImplementedError):  # prefill stays the reference's (no delegate registered here)
        fbatch.forward_mode = fb.ForwardMode.EXTEND
        backend.forward_extend(q, k_new, k_new, layer, fbatch)


def test_multi_step_backend_matches_reference_signature_and_lengths():
    """FlashMLAMultiStepDecodeBackend as speculative/base_spec_worker.py:238-240 constructs it --
    ``(draft_model_runner, speculative_num_steps)`` -- with step i attending over seq_lens + i + 1
    (flashmla_backend.py:411-422): three draft steps, each writes the K row of its own token and then attends; every
    step is compared with fp32 attention over ctx + i + 1 rows."""
    heads, steps = 16, 3
    fb, mr, alloc, pool, layer = _runner(heads)
    ms = fb.FlashMLAMultiStepDecodeBackend(mr, steps)
    assert [b.multi_step_seqlen_incr for b in ms.attn_backends] == [1, 2, 3]
    assert ms.get_cuda_graph_seq_len_fill_value() == 1
    with pytest.raises(TypeError):  # the round-1 signature (model_runner, topk, steps) is gone
        fb.FlashMLAMultiStepDecodeBackend(mr, 1, steps, 1)
    ctx = [130, 64, 700, 1]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    seq_lens = torch.tensor(ctx, device=DEV)          # the verified context: constant over the draft steps
    locs = [l.clone() for l in ctx_locs]
    torch.manual_seed(11)
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=seq_lens, req_pool_indices=torch.arange(bs, device=DEV),
                             out_cache_loc=None, token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DECODE,
                             spec_info=object())
    # the scheduler reserves the slots of all draft tokens before the draft loop starts (the block table the metadata
    # snapshots must already hold their pages)
    draft_locs = [alloc.alloc(r, steps, n) for r, n in enumerate(ctx)]
    ms.init_forward_metadata(fbatch)                   # common_template: the first steps - 1 backends
    assert ms.attn_backends[0].forward_metadata is not None and ms.attn_backends[steps - 1].forward_metadata is None
    ms.attn_backends[steps - 1].init_forward_metadata(fbatch)
    for i in range(steps):
        new_loc = torch.stack([draft_locs[r][i] for r in range(bs)])
        fbatch.out_cache_loc = new_loc
        q = torch.randn(bs, heads * 576, dtype=torch.bfloat16, device=DEV)
        k_new = torch.randn(bs, 1, 576, dtype=torch.bfloat16, device=DEV)
        out = ms.attn_backends[i].forward_decode(q, k_new, k_new[..., :512], layer, fbatch)
        locs = [torch.cat([locs[r], new_loc[r:r + 1]]) for r in range(bs)]
        ref = _reference(pool, layer, q.view(bs, 1, heads, 576), locs, 1, False)   # ctx + i + 1 rows per request
        assert _rel(out.view(bs, 1, heads, 512), ref) < 5e-2, i


def test_backend_resolves_attention_tp_like_the_reference():
    """attn_initializer.py:395 constructs FlashMLABackend(model_runner) with no TP argument; heads per rank come from
    get_attention_tp_size() (flashmla_backend.py:67-69).  128 heads at attention-TP 8 -> 16 local heads."""
    heads, tp = 128, 8
    fb, mr, alloc, pool, layer = _runner(heads)
    fb.set_attention_tp_size_getter(lambda: tp)
    try:
        backend = fb.FlashMLABackend(mr)
    finally:
        fb.set_attention_tp_size_getter(None)
    assert backend.num_q_heads == backend.num_local_heads == heads // tp
    local = SimpleNamespace(layer_id=1, tp_q_head_num=heads // tp, head_dim=576, v_head_dim=512, scaling=SCALING)
    ctx = [200, 65]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, local, ctx)
    new_loc = torch.cat([alloc.alloc(i, 1, n) for i, n in enumerate(ctx)])
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=torch.tensor([n + 1 for n in ctx], device=DEV),
                             req_pool_indices=torch.arange(bs, device=DEV), out_cache_loc=new_loc,
                             token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DECODE)
    backend.init_forward_metadata(fbatch)
    torch.manual_seed(5)
    q = torch.randn(bs, (heads // tp) * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs, 1, 576, dtype=torch.bfloat16, device=DEV)
    out = backend.forward(q, k_new, k_new[..., :512], local, fbatch)
    all_locs = [torch.cat([ctx_locs[i], new_loc[i:i + 1]]) for i in range(bs)]
    ref = _reference(pool, local, q.view(bs, 1, heads // tp, 576), all_locs, 1, False)
    assert _rel(out.view(bs, 1, heads // tp, 512), ref) < 5e-2


def test_backend_draft_extend_and_idle():
    """DRAFT_EXTEND runs the verify kernel shape (flashmla_backend.py:105-176); an IDLE batch (a data-parallel attention
    rank without requests: batch_size 0, empty seq_lens) gets metadata and an empty result instead of an error."""
    heads, draft = 16, 4
    fb, mr, alloc, pool, layer = _runner(heads, draft_tokens=draft)
    ctx = [150, 64]
    bs = len(ctx)
    ctx_locs = _fill_context(alloc, pool, layer, ctx)
    backend = fb.FlashMLABackend(mr)
    new_loc = torch.cat([alloc.alloc(i, draft, n) for i, n in enumerate(ctx)])
    fbatch = SimpleNamespace(batch_size=bs, seq_lens=torch.tensor(ctx, device=DEV), req_pool_indices=torch.arange(bs, device=DEV),
                             out_cache_loc=new_loc, token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DRAFT_EXTEND)
    backend.init_forward_metadata(fbatch)
    torch.manual_seed(6)
    q = torch.randn(bs * draft, heads * 576, dtype=torch.bfloat16, device=DEV)
    k_new = torch.randn(bs * draft, 1, 576, dtype=torch.bfloat16, device=DEV)
    out = backend.forward(q, k_new, k_new[..., :512], layer, fbatch)
    all_locs = [torch.cat([ctx_locs[i], new_loc[i * draft:(i + 1) * draft]]) for i in range(bs)]
    ref = _reference(pool, layer, q.view(bs, draft, heads, 576), all_locs, draft, True)
    assert _rel(out.view(bs, draft, heads, 512), ref) < 5e-2
    # IDLE
    fb2, mr2, alloc2, pool2, layer2 = _runner(heads)
    idle_backend = fb2.FlashMLABackend(mr2)
    idle = SimpleNamespace(batch_size=0, seq_lens=torch.empty(0, dtype=torch.int64, device=DEV),
                           req_pool_indices=torch.empty(0, dtype=torch.int64, device=DEV),
                           out_cache_loc=torch.empty(0, dtype=torch.int32, device=DEV), token_to_kv_pool=pool2,
                           forward_mode=fb2.ForwardMode.IDLE)
    idle_backend.init_forward_metadata(idle)
    meta = idle_backend.forward_metadata
    assert meta.num_splits.cpu().tolist() == [0] and meta.block_table.shape[0] == 0
    assert (meta.flashmla_metadata[:, 0] >= 0).all()          # every part starts at request 0 >= batch: no work
    out = idle_backend.forward_decode(torch.empty(0, heads * 576, dtype=torch.bfloat16, device=DEV), None, None, layer2, idle)
    assert out.shape == (0, heads * 512)
    torch.cuda.synchronize()


def test_backend_prefill_is_delegated_when_a_prefill_backend_is_registered():
    """EXTEND / MIXED go to the reference's FlashInferMLAAttnBackend (flashmla_backend.py:100-104, 289-290): here a
    registered delegate built with the constructor's arguments; without one the call raises."""
    heads = 16
    fb, mr, alloc, pool, layer = _runner(heads)
    calls = []

    class Prefill:
        def __init__(self, *a):
            calls.append(("init", a))

        def init_forward_metadata(self, forward_batch):
            calls.append(("meta", forward_batch.forward_mode))

        def forward_extend(self, q, k, v, layer, forward_batch, save_kv_cache, q_pe, k_pe):
            calls.append(("extend", save_kv_cache))
            return "prefill-result"

    fbatch = SimpleNamespace(batch_size=1, forward_mode=fb.ForwardMode.EXTEND)
    plain = fb.FlashMLABackend(mr)
    with pytest.raises(NotImplementedError):
        plain.forward_extend(None, None, None, layer, fbatch)
    with pytest.raises(RuntimeError):
        plain.init_forward_metadata(fbatch)
    fb.set_prefill_backend_factory(lambda *a: Prefill(*a))
    try:
        backend = fb.FlashMLABackend(mr, False, "kv_indptr", "q_indptr")
        assert fb.FlashMLABackend(mr, skip_prefill=True).prefill_backend is None
    finally:
        fb.set_prefill_backend_factory(None)
    assert calls[0] == ("init", (mr, False, "kv_indptr", "q_indptr"))
    backend.init_forward_metadata(fbatch)
    assert backend.forward(None, None, None, layer, fbatch) == "prefill-result"
    fbatch.forward_mode = fb.ForwardMode.MIXED
    backend.init_forward_metadata(fbatch)
    assert [c[0] for c in calls] == ["init", "meta", "extend", "meta"]


@pytest.mark.parametrize("m,n,k", [(5, 192, 256), (128, 2112, 1536)])
def test_fp8_linear_method_apply(m, n, k):
    """Fp8LinearMethod.apply: 1x128 activation quantisation + block-FP8 GEMM + bias, against the oracle."""
    from sglang_fluentllm_b200 import deep_gemm as dg
    from sglang_fluentllm_b200 import fp8_linear as fl
    torch.manual_seed(4)
    method = fl.Fp8LinearMethod(fl.Fp8Config(True, "dynamic", None, [128, 128]))
    layer = torch.nn.Module()
    method.create_weights(layer, k, [n], k, n, torch.bfloat16, device=DEV)
    w = (torch.randn(n, k, device=DEV) * k ** -0.5).to(torch.bfloat16)
    w8, ws = quantize_rhs_128x128(w)
    layer.weight.data.copy_(w8)
    layer.weight_scale_inv.data.copy_(ws)
    method.process_weights_after_loading(layer)
    x = torch.randn(2, m, k, dtype=torch.bfloat16, device=DEV)
    bias = torch.randn(n, dtype=torch.bfloat16, device=DEV)
    y = method.apply(layer, x, bias=bias)
    assert y.shape == (2, m, n) and y.dtype == torch.bfloat16
    xq, xs = ogemm.per_token_group_quant_fp8(x.view(-1, k).float().cpu().numpy())
    ref = ogemm.w8a8_block_fp8_matmul(ofp8.e4m3_to_f32(xq), xs, ofp8.e4m3_to_f32(w8.view(torch.uint8).cpu().numpy()),
                                      ws.cpu().numpy()) + bias.float().cpu().numpy()
    got = y.view(-1, n).float().cpu().numpy()
    assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 5e-3
    with pytest.raises(RuntimeError, match="no fallback"):
        method.apply(layer, x, output_dtype=torch.float16)
