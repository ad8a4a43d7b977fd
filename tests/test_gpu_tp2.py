"""Two GPUs, NCCL: the tensor-parallel partition of the attention block (heads / column-parallel q_b_proj / row-parallel
o_proj + ONE all-reduce, sglang_fluentllm_b200/tp.py) computed with the CUDA kernels must equal the unsharded block to
bf16 tolerance (reference: python/sglang/srt/layers/linear.py:199,963,1139, flashmla_backend.py:67-75).
Runs under `gpurun --gpus 2`; skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu

HEADS, D_QK, D_V, Q_LORA, HIDDEN, B, SEQ = 32, 192, 128, 1536, 2048, 24, 700


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(dev, plan_rank, plan_size):
    """Everything is generated from fixed seeds on every rank, so the ranks (and the unsharded run) hold identical
    weights, pool and inputs."""
    from sglang_fluentllm_b200 import deep_gemm as dg, flash_mla_fp8 as fm, fp8_kernel as fk, tp
    from sglang_fluentllm_b200.fp8_linear import dummy_block_fp8_weight
    g = torch.Generator(device=dev).manual_seed(1234)
    w_q_b = dummy_block_fp8_weight(HEADS * D_QK, Q_LORA, dev, g)
    w_o = dummy_block_fp8_weight(HIDDEN, HEADS * D_V, dev, g)
    w_kc = (torch.randn(HEADS, 128, 512, device=dev, generator=g) * 128 ** -0.5).to(torch.bfloat16)   # absorb: q_nope -> latent
    w_vc = (torch.randn(HEADS, 512, D_V, device=dev, generator=g) * 512 ** -0.5).to(torch.bfloat16)   # latent out -> v head
    q_a = torch.randn(B, Q_LORA, device=dev, generator=g).to(torch.bfloat16)
    pages = (SEQ + 63) // 64 + 1
    n_pages = B * pages + 1
    kv = torch.randn(n_pages * 64, 1, 576, device=dev, generator=g).to(torch.bfloat16)
    kl = torch.zeros(n_pages * 64, 1, 512, dtype=torch.uint8, device=dev)
    ks = torch.zeros(n_pages * 64, 1, 1, dtype=torch.float32, device=dev)
    kr = torch.zeros(n_pages * 64, 1, 64, dtype=torch.bfloat16, device=dev)
    fm.quantize_and_cache_k(kv, kl, ks, kr, torch.arange(n_pages * 64, dtype=torch.int32, device=dev), 512)
    bt = (torch.arange(B * pages, device=dev).reshape(B, pages) + 1).to(torch.int32)
    sl = torch.full((B,), SEQ, dtype=torch.int32, device=dev)
    plan = tp.TpPlan(rank=plan_rank, size=plan_size, num_heads=HEADS, qk_head_dim=D_QK, v_head_dim=D_V)
    h0, h1 = plan.head_range

    def linear(x, ws):
        w, s = ws
        xq, xs = fk.sglang_per_token_group_quant_fp8(x.contiguous(), 128, column_major_scales=True, scale_tma_aligned=True)
        y = torch.empty(x.shape[0], w.shape[0], dtype=torch.bfloat16, device=dev)
        dg.gemm_fp8_fp8_bf16_nt((xq, xs), (w, s), y, False)
        return y

    def attention(q):  # [b, h_local, 192] -> [b, h_local, 128] over this rank's heads, full KV pool on every rank
        hl = q.shape[1]
        q_lat = torch.einsum("bhd,hdc->bhc", q[..., :128], w_kc[h0:h1])              # w_kc absorb (deepseek_v2.py:840)
        q_abs = torch.cat([q_lat, q[..., 128:]], dim=-1).unsqueeze(1).contiguous()   # [b, 1, h_local, 576]
        qn, qs, qr = fm.quantize_ckv_per_token_head(q_abs, 512)
        meta, ns = fm.get_mla_metadata(sl, hl, 1)
        o, _ = fm.flash_mla_ckv_fp8_per_token(qn, qr, kl.view(n_pages, 64, 1, 512), kr.view(n_pages, 64, 1, 64), qs,
                                              ks.view(n_pages, 64, 1, 1), bt, sl, 512, meta, ns, softmax_scale=192 ** -0.5)
        return torch.einsum("bhc,hcd->bhd", o[:, 0], w_vc[h0:h1])                     # w_vc absorb (deepseek_v2.py:886)

    return tp.TensorParallelAttentionBlock(plan, w_q_b, w_o, linear, attention), q_a


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        block, q_a = _build(dev, rank, world)
        y = block.forward(q_a)                      # all-reduced: identical on both ranks
        torch.cuda.synchronize()
        gathered = [torch.empty_like(y) for _ in range(world)]
        dist.all_gather(gathered, y)
        if rank == 0:
            assert torch.equal(gathered[0], gathered[1])
            single, q_a1 = _build(dev, 0, 1)        # the unsharded block on one GPU (no process-group exchange: size 1)
            single.group = None
            p = single.plan
            q = single.linear(q_a1, single.w_q_b)
            o = single.attention(q.view(B, HEADS, D_QK))
            ref = single.linear(o.reshape(B, HEADS * D_V), single.w_o)
            torch.cuda.synchronize()
            np.savez(out_path, y=y.float().cpu().numpy(), ref=ref.float().cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_tp2_block_equals_unsharded_block(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    out_path = str(tmp_path / "tp2.npz")
    mp.spawn(_worker, args=(2, _free_port(), out_path), nprocs=2, join=True)
    r = np.load(out_path)
    rel = np.linalg.norm(r["y"] - r["ref"]) / np.linalg.norm(r["ref"])
    # the sharded run rounds each rank's partial o_proj sums to bf16 before the all-reduce and re-quantises per-rank
    # activations: bf16-level agreement
    assert rel < 1e-2, rel


def _p2p_worker(rank, world, port, out_path):
    """The peer-memory collectives of csrc/p2p_comm.cu against torch.distributed (NCCL) on the same tensors."""
    import torch.distributed as dist
    from sglang_fluentllm_b200 import comm
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ.setdefault("B200_COMM_SPIN_LIMIT", str(1 << 24))  # a lost signal fails within seconds here
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", rank)
        hidden = 7168
        rsag = comm.RSAG(dist.group.WORLD, rank, max_tokens=512, hidden_size=hidden, device=dev)
        p2p = rsag.p2p
        g = torch.Generator(device=dev).manual_seed(7 + rank)
        res = {}

        def same_on_every_rank(t):
            parts = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(parts, t.contiguous())
            return all(torch.equal(parts[0], q) for q in parts[1:])

        def close(got, want_f32):
            # two ranks: one fp32 add, the same bits as NCCL's fp32 sum rounded to bf16; more ranks: the order of the
            # fp32 adds differs (rank order here), so the bf16 results may differ by a rounding at a tie
            if world == 2:
                return torch.equal(got, want_f32.to(torch.bfloat16))
            return torch.allclose(got.float(), want_f32, rtol=2 ** -7, atol=1e-6)

        # all-reduce, 128 x 7168 (the o_proj output of the metric shape) and small ragged sizes; repeated (the buffer
        # halves alternate, the call counters advance)
        for shape in ((128, hidden), (3, 8), (37, 1024)):
            for it in range(3):
                x = torch.randn(*shape, device=dev, generator=g).to(torch.bfloat16)
                want = x.float().clone()
                dist.all_reduce(want)
                got = p2p.all_reduce(x.clone())
                torch.cuda.synchronize()
                assert close(got, want), (shape, it)
                assert same_on_every_rank(got), (shape, it)
        # uneven reduce-scatter / all-gather, incl. ranks without rows
        lists = [[40 - r for r in range(world)], [5] + [0] * (world - 1), [0] * (world - 1) + [3], [1] * world,
                 comm.token_dist(37, world)]
        for tokens in lists:
            total = sum(tokens)
            x = torch.randn(total, hidden, device=dev, generator=g).to(torch.bfloat16)
            want = x.float().clone()
            dist.all_reduce(want)
            mine, off = rsag.reduce_scatter(x, token_list_in_group=tokens)
            assert off == sum(tokens[:rank]) and mine.shape[0] == tokens[rank]
            assert close(mine, want[off: off + tokens[rank]]), tokens
            back = rsag.all_gather(mine, token_list_in_group=tokens)
            torch.cuda.synchronize()
            assert close(back, want), tokens
            assert same_on_every_rank(back), tokens
        mine, off = rsag.reduce_scatter(torch.ones(world + 1, hidden, device=dev, dtype=torch.bfloat16), world + 1)
        assert (mine.shape[0], off) == ((2, 0) if rank == 0 else (1, rank + 1)) and float(mine.float().min()) == world
        # the all-reduce inside a CUDA graph, replayed (the call counters live in device memory)
        x = torch.randn(128, hidden, device=dev, generator=g).to(torch.bfloat16)
        y = torch.empty_like(x)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            p2p.all_reduce(x, out=y)
            s.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=s):
                p2p.all_reduce(x, out=y)
            want = x.float().clone()
        dist.all_reduce(want)
        for _ in range(5):
            y.zero_()
            graph.replay()
            torch.cuda.synchronize()
            assert close(y, want)

        # timing, 1.84 MB: NCCL against the peer-memory kernel (device-timed, max over ranks)
        def timed(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1) * 1e3 / n], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        res["world"] = world
        res["mode"] = os.environ.get("B200_COMM_ONE_SHOT", "default")
        res["nccl_us"] = timed(lambda: dist.all_reduce(x))
        res["p2p_us"] = timed(lambda: p2p.all_reduce(x, out=y, nvls=False))
        res["p2p_graph_us"] = timed(graph.replay)
        # the same sum through the switch, where the box has NVSwitch multicast
        res["nvls"] = p2p.enable_nvls()
        res["nvls_error"] = p2p.nvls_error
        if res["nvls"]:
            for shape in ((128, hidden), (3, 8), (37, 1024)):
                for it in range(3):
                    xs = torch.randn(*shape, device=dev, generator=g).to(torch.bfloat16)
                    want = xs.float().clone()
                    dist.all_reduce(want)
                    got = p2p.all_reduce(xs.clone(), nvls=True)
                    torch.cuda.synchronize()
                    assert torch.allclose(got.float(), want, rtol=2 ** -7, atol=1e-6), (shape, it)
                    assert same_on_every_rank(got), (shape, it)
            # interleaved with the peer-memory calls (they share the call counters and alternate the buffer halves)
            for it in range(4):
                xs = torch.randn(64, hidden, device=dev, generator=g).to(torch.bfloat16)
                want = xs.float().clone()
                dist.all_reduce(want)
                got = p2p.all_reduce(xs.clone(), nvls=bool(it & 1))
                torch.cuda.synchronize()
                assert torch.allclose(got.float(), want, rtol=2 ** -7, atol=1e-6), it
            res["nvls_us"] = timed(lambda: p2p.all_reduce(x, out=y, nvls=True))
        if rank == 0:
            import json
            json.dump(res, open(out_path, "w"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("one_shot", ["default", "0"], ids=["default", "two-shot"])
def test_p2p_collectives_match_nccl(tmp_path, one_shot, monkeypatch):
    """On every GPU of the box (2, 4 or 8): the group sizes of one NVLink node.  'two-shot' forces the reduce-a-slice /
    collect form of the all-reduce that groups of more than two ranks take by default."""
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs (gpurun --gpus 2)")
    world = 8 if world >= 8 else (4 if world >= 4 else 2)
    if one_shot != "default":
        monkeypatch.setenv("B200_COMM_ONE_SHOT", one_shot)
    out_path = str(tmp_path / "p2p.json")
    mp.spawn(_p2p_worker, args=(world, _free_port(), out_path), nprocs=world, join=True)
    import json
    r = json.load(open(out_path))
    print("all-reduce of 128 x 7168 bf16 on %d GPUs (%s): NCCL %.1f us, peer-memory kernel %.1f us (in a graph %.1f us), "
          "through the switch %s" % (world, one_shot, r["nccl_us"], r["p2p_us"], r["p2p_graph_us"],
                                      ("%.1f us" % r["nvls_us"]) if r.get("nvls") else "unavailable: %s" % r.get("nvls_error")))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(r, open("gpurun_out/p2p_all_reduce_%dgpu_%s.json" % (world, one_shot), "w"))
