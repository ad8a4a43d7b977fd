"""CPU: the exchanges either side of the attention block (sglang_fluentllm_b200/comm.py) -- token distribution, the
decision table of DecoderCommManager against the REFERENCE's own class imported from /root/reference (every strategy
combination, recorded calls and results), and RSAG's uneven reduce-scatter / all-gather semantics over gloo, world 2."""
import importlib.util
import itertools
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sglang_fluentllm_b200 import comm

REFERENCE_PY = "/root/reference/python"


def test_token_dist_matches_reference_rule():
    # triton_rsag.py:48-57
    assert comm.token_dist(10, 4) == [3, 3, 2, 2]
    assert comm.token_dist(3, 8) == [1, 1, 1, 0, 0, 0, 0, 0]
    assert comm.token_dist(0, 2) == [0, 0]
    assert sum(comm.token_dist(12345, 7)) == 12345


class _FakeRsag:
    """Records calls; 'reduce-scatter' keeps rows [2, 5), 'all-gather' doubles the rows."""

    def __init__(self, log):
        self.log = log

    def reduce_scatter(self, h, n, *a, **k):
        self.log.append(("rs", n))
        return h[2:5] + 1, 2

    def all_gather(self, h, n, *a, **k):
        self.log.append(("ag", n))
        return torch.cat([h, h])


def _run(mgr, log):
    """Drive one manager through the three attention-side hooks; returns what it did (or the exception type)."""
    h = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4)
    res = torch.arange(8 * 4, dtype=torch.float32).reshape(8, 4) * 10
    out = []
    for name, call in (("pre", lambda: mgr.pre_attn_comm(h, 8)), ("pre2", lambda: mgr.pre_attn_comm(h, 8, True)),
                       ("post", lambda: mgr.post_attn_comm(h, res, 8)), ("post_nf", lambda: mgr.post_attn_comm(h, res, 8, False)),
                       ("final", lambda: mgr.post_final_norm_comm(h, res, 8))):
        del log[:]
        try:
            r = call()
            r = r if isinstance(r, tuple) else (r,)
            out.append((name, tuple(log), tuple(t.shape for t in r), tuple(float(t.sum()) for t in r)))
        except NotImplementedError:
            out.append((name, "NotImplementedError"))
    return out


@pytest.mark.skipif(not os.path.isdir(REFERENCE_PY), reason="needs the reference tree")
def test_decision_table_matches_the_references_own_class(monkeypatch):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    finder = mg._Finder()
    sys.meta_path.insert(0, finder)
    sys.path.insert(0, REFERENCE_PY)
    before = set(sys.modules)
    try:
        import sglang.srt.distributed.decoder_comm_manager as ref_mod
        from sglang.srt.distributed.parallel_strategy import AttnParallelStrategy, DenseParallelStategy, MoeParallelStrategy
        ref_log, our_log = [], []
        monkeypatch.setattr(ref_mod, "tensor_model_parallel_all_reduce", lambda t: (ref_log.append(("ar",)), t * 2)[1])
        monkeypatch.setattr(ref_mod, "FLLM_IS_CP", False)
        n = 0
        for attn, dense, moe, is_moe, layer_id, sizes in itertools.product(
                AttnParallelStrategy, DenseParallelStategy, MoeParallelStrategy, (False, True), (0, 3),
                ((8, 8, 8), (4, 8, 8), (4, 4, 8), (1, 8, 8))):
            attn_tp, dense_tp, tp = sizes
            ref = object.__new__(ref_mod.DecoderCommMananger)
            ref.attn_parallel_strategy, ref.dense_parallel_strategy, ref.moe_parallel_strategy = attn, dense, moe
            ref.is_moe_layer, ref.layer_id, ref.num_layers = is_moe, layer_id, 4
            ref.attn_tp_size, ref.dense_tp_size, ref.tp_size = attn_tp, dense_tp, tp
            ref.attn_rsag = ref.dense_rsag = _FakeRsag(ref_log)
            ours = comm.DecoderCommManager(layer_id, attn.value, dense.value, moe.value, is_moe, 4, attn_tp_size=attn_tp,
                                           dense_tp_size=dense_tp, tp_size=tp, attn_rsag=_FakeRsag(our_log),
                                           all_reduce=lambda t: (our_log.append(("ar",)), t * 2)[1])
            if attn_tp == 1:
                ours.attn_rsag = _FakeRsag(our_log)  # the reference object above keeps one regardless of the size
            assert _run(ours, our_log) == _run(ref, ref_log), (attn, dense, moe, is_moe, layer_id, sizes)
            n += 1
        assert n == 3 * 3 * 2 * 2 * 2 * 4
    finally:
        sys.meta_path.remove(finder)
        sys.path.remove(REFERENCE_PY)
        for m in set(sys.modules) - before:
            if m.split(".")[0] in ("sglang",) or m.split(".")[0] in mg.STUBS:
                sys.modules.pop(m, None)


def test_decision_table_fixed_points():
    """A few rows of the table spelled out (decoder_comm_manager.py:152-262), for machines without the reference tree."""
    log = []
    mk = lambda attn, dense, moe, is_moe, layer, sizes=(8, 8, 8): comm.DecoderCommManager(
        layer, attn, dense, moe, is_moe, 4, attn_tp_size=sizes[0], dense_tp_size=sizes[1], tp_size=sizes[2],
        attn_rsag=_FakeRsag(log), all_reduce=lambda t: (log.append(("ar",)), t)[1])
    h, res = torch.zeros(8, 4), torch.ones(8, 4)
    # TP attention + TP MoE: one all-reduce behind the block, nothing in front (:191-196)
    m = mk("tp", "tp", "tp", True, 1)
    m.post_attn_comm(h, res, 8)
    assert log == [("ar",)] and m.pre_attn_comm(h, 8) is h
    # TP attention + expert-parallel MoE: reduce-scatter behind, all-gather in front from the second block on; the
    # first residual is cut to the rank's rows (:197-205, :156-163)
    del log[:]
    m = mk("tp", "tp", "ep", True, 0)
    out, r = m.post_attn_comm(h, res, 8)
    assert log == [("rs", 8)] and out.shape[0] == 3 and r.shape[0] == 3
    assert m.pre_attn_comm(h, 8) is h and m.pre_attn_comm(h, 8, is_second_attn=True).shape[0] == 16
    # DP attention: no exchange around the block (:206-217)
    del log[:]
    m = mk("dp", "rep", "ep", False, 2, (1, 1, 8))
    assert m.post_attn_comm(h, res, 8) == (h, res) and m.pre_attn_comm(h, 8) is h and log == []
    with pytest.raises(NotImplementedError):
        mk("combine", "tp", "tp", True, 1).post_attn_comm(h, res, 8)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rsag = comm.RSAG(dist.group.WORLD, rank, max_tokens=16, hidden_size=8)
        g = torch.Generator().manual_seed(100 + rank)
        total = 7                                       # uneven: 4 rows on rank 0, 3 on rank 1
        x = torch.randn(total, 8, generator=g).to(torch.bfloat16)
        mine, off = rsag.reduce_scatter(x, total)
        back = rsag.all_gather(mine, total)
        tokens = [2, 5]                                 # an explicit token list (decoder_comm_manager.py:343-344)
        mine2, off2 = rsag.reduce_scatter(x, token_list_in_group=tokens)
        # TP attention + replicated dense through the manager: scatter behind block 0, gather in front of block 1
        mgr = comm.DecoderCommManager(0, "tp", "rep", "tp", False, 2, attn_tp_size=2, dense_tp_size=1, tp_size=2,
                                      attn_rsag=rsag, all_reduce=None)
        h, res = mgr.post_attn_comm(x, x.clone(), total)
        mgr1 = comm.DecoderCommManager(1, "tp", "rep", "tp", False, 2, attn_tp_size=2, dense_tp_size=1, tp_size=2,
                                       attn_rsag=rsag, all_reduce=None)
        full = mgr1.pre_attn_comm(h, total)
        out[rank] = dict(x=x, mine=mine, off=off, back=back, mine2=mine2, off2=off2, h=h, res=res, full=full)
    finally:
        dist.destroy_process_group()


def test_rsag_uneven_reduce_scatter_all_gather_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    total = (out[0]["x"].float() + out[1]["x"].float()).to(torch.bfloat16)
    assert (out[0]["off"], out[1]["off"]) == (0, 4)
    assert torch.equal(out[0]["mine"], total[:4]) and torch.equal(out[1]["mine"], total[4:])
    for r in range(world):
        assert torch.equal(out[r]["back"], total)
        assert torch.equal(out[r]["full"], total)
    assert (out[0]["off2"], out[1]["off2"]) == (0, 2)
    assert torch.equal(out[0]["mine2"], total[:2]) and torch.equal(out[1]["mine2"], total[2:])
    assert torch.equal(out[0]["res"], out[0]["x"][:4]) and torch.equal(out[1]["res"], out[1]["x"][4:])
