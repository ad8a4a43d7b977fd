#!/usr/bin/env python
"""bench.py -- decode tokens/s of the quantised decode hot path (DeepSeek-V3 MLA, per-token-FP8 KV, bs=128, seq=4K).

One "step" = one decode step of the attention block (models/deepseek_v2.py forward_absorb, :796-889) through all 61
layers for one batch, driven through the plug-in API the reference's model runner uses: per layer
  Fp8LinearMethod.apply (fused_qkv_a_proj: activation quantiser + block-FP8 GEMM) -> FusedRMSNorm of q_a / kv_a with
  the q_b_proj input emitted block-quantised -> Fp8LinearMethod.apply (q_b_proj, block_scale supplied) -> w_kc absorb
  (bf16 torch.bmm, the reference's own op) -> rotary embedding with the fused set_kv_buffer argument (rope of q_pe /
  k_pe, Q quantiser, quantise-and-cache of the new token's K: one launch) -> FlashMLABackend.forward_decode (paged MLA
  decode over the FP8 KV pool + split-KV combine) -> w_vc absorb (bmm) -> strided activation quantiser ->
  Fp8LinearMethod.apply (o_proj, row-parallel) -> [all-reduce when tp > 1].
The backend is used exactly as cuda_graph_runner.py:405-526 uses it: init_cuda_graph_state, capture-time metadata,
the layer loop captured in a CUDA graph, and init_forward_metadata_replay_cuda_graph outside the graph before a replay.

tokens/s = batch / step time.  `value` is timed with inputs resident in HBM (CUDA-graph replay of the step);
`e2e` adds, inside the timed region, the pinned-host -> device copies of the step inputs (hidden states, sequence
lengths, request slots, positions, cache slots, page-table rows), the backend's replay-metadata refresh (block-table
gather + tile-scheduler kernel) and a device -> host read of the step output.

Multi-GPU (torchrun, one rank per GPU).  Default `--parallelism dp`: one replica of the path per GPU with the per-GPU
batch fixed at 128 (weak scaling, no data-path collective) -- MLA has a single latent KV head, so under head tensor
parallelism every rank streams the whole KV pool and attention time does not shrink; sharding the
requests is what the reference's DP-attention does to this block.  The same run then measures strict head-TP
(`--parallelism tp`: heads / column-parallel q_b_proj / row-parallel o_proj over the GPUs, global batch fixed, ONE NCCL
all-reduce per layer on the row-parallel output over NVLink) and reports it under "tp_strong".

`--impl reference` times the reference's CPU path (torch-native SDPA decode loop + native block-FP8 matmul,
oracle/native_cpu.py -- a port, the reference package is not installable) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

HIDDEN, Q_LORA, KV_LORA, ROPE, NOPE_HEAD, V_HEAD, HEADS, LAYERS = 7168, 1536, 512, 64, 128, 128, 128, 61
PAGE = 64
METRIC = "decode tokens/s/GPU DeepSeek-V3 MLA FP8-KV bs=128 seq=4K"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--seq", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=LAYERS)
    ap.add_argument("--heads", type=int, default=HEADS)
    ap.add_argument("--parallelism", default=None, choices=["tp", "dp"],
                    help="N > 1: dp (default) = one replica of the path per GPU, per-GPU batch fixed (weak scaling, no "
                         "data-path collective; what the reference's DP-attention does to this block); tp = heads / "
                         "column- and row-parallel weights over the GPUs with one all-reduce per layer (strong scaling; "
                         "MLA's single latent KV head makes every rank stream the whole KV pool).  The dp run also "
                         "reports a short tp measurement under \"tp_strong\".")
    ap.add_argument("--timeline", default=None, metavar="FILE",
                    help="diagnostic: replay the captured step under the CUPTI activity profiler and write per-kernel "
                         "durations and the idle gaps between consecutive kernels to FILE (not a bench value)")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop and exit (for an ncu launch list of the step)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML, ~2 ms period; nvidia-smi as a fallback)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.sm, self.power = [], []
        self.sm_max = None
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self._nvml = pynvml
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.sm.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
        try:
            self.power.append(n.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
        except Exception:
            pass
        r = int(n.nvmlDeviceGetCurrentClocksEventReasons(self._h)) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
        for name, bit in (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20),
                          ("sw_power_cap", 0x4)):
            if r & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                              "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
        parts = [p.strip() for p in out.strip().split(",")]
        if len(parts) >= 7:
            self.sm.append(int(float(parts[0])))
            self.sm_max = int(float(parts[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def _run(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml() if self._nvml else self._sample_smi()
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=6)
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_min_mhz": sm[0] if sm else None,
                "sm_max_mhz": self.sm_max, "power_w_max": round(max(self.power), 1) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# the B200 arm
# ----------------------------------------------------------------------------------------------------------------
class HotPath:
    """The benchmarked step for one rank, built the way the reference's model runner builds it around the plug-in:
    a KVAllocator + MLATokenToKVPool (per-token-FP8 layout), ONE FlashMLABackend driven through the CUDA-graph protocol
    of cuda_graph_runner.py:405-526 (init_cuda_graph_state -> ...capture_cuda_graph -> graph capture of the layer loop ->
    per step ...replay_cuda_graph outside the graph + replay), Fp8LinearMethod.apply for the three block-FP8 linears,
    FusedRMSNorm for q_a / kv_a, the rotary-embedding prologue with the fused set_kv_buffer argument, and the w_kc / w_vc
    absorb bmm's (models/deepseek_v2.py:796-889, forward_absorb).  No stand-ins: the step is a runnable attention block."""

    def __init__(self, args, rank, world, dev):
        import torch
        from types import SimpleNamespace

        from sglang_fluentllm_b200 import flash_mla_fp8 as fm
        from sglang_fluentllm_b200 import flashmla_backend as fb
        from sglang_fluentllm_b200 import fp8_kernel as fk
        from sglang_fluentllm_b200 import fp8_linear as fl
        from sglang_fluentllm_b200 import layernorm as ln
        from sglang_fluentllm_b200 import rope as rope_mod
        from sglang_fluentllm_b200 import tp as tpmod
        self.torch, self.args, self.dev, self.world = torch, args, dev, world
        self.fm, self.fb, self.fk, self.rope_mod, self.tpmod = fm, fb, fk, rope_mod, tpmod
        tp = world if args.parallelism == "tp" else 1
        self.tp = tp
        b, seq, L = args.batch, args.seq, args.layers
        # heads / column-parallel q_b_proj rows / row-parallel o_proj columns of this rank (sglang_fluentllm_b200/tp.py)
        self.plan = tpmod.TpPlan(rank if tp > 1 else 0, tp, args.heads, NOPE_HEAD + ROPE, V_HEAD)
        self.h = h = self.plan.heads_per_rank
        g = torch.Generator(device=dev)
        g.manual_seed(1234 + rank)
        pages_per_req = (seq + PAGE - 1) // PAGE + 1
        self.n_pages = b * pages_per_req + 1
        slots = self.n_pages * PAGE
        scaling = (NOPE_HEAD + ROPE) ** -0.5

        # ---- the plug-in's surroundings: allocator (page table), pool, model-runner facade ----
        ctx_len = pages_per_req * PAGE
        self.alloc = fb.KVAllocator(size=slots, device=dev, max_batch_size=b, max_context_len=ctx_len)
        # page table of a long-running server: every request owns `pages_per_req` shuffled physical pages (page 0 is the
        # reserved padding page); the host copy is what the scheduler writes through on allocation
        perm = (torch.randperm(self.n_pages - 1, generator=torch.Generator().manual_seed(7)) + 1).to(torch.int32)
        self.page_table_host = perm[: b * pages_per_req].reshape(b, pages_per_req).contiguous().pin_memory()
        self.alloc.req_to_page[:b, :pages_per_req].copy_(self.page_table_host)
        self.pool = fb.MLATokenToKVPool(size=slots - PAGE, model_dtype=torch.bfloat16, kv_lora_rank=KV_LORA,
                                        qk_rope_head_dim=ROPE, layer_num=L, device=dev, quant_method="per_token_head")
        tmp = torch.empty(slots, 1, KV_LORA + ROPE, dtype=torch.bfloat16, device=dev)
        idx = torch.arange(slots, dtype=torch.int32, device=dev)
        for l in range(L):  # N(0, 1) context, quantised by the path's own quantiser
            tmp.normal_(generator=g)
            kl, ks, kr = self.pool.get_key_buffer(l)
            fm.quantize_and_cache_k(tmp, kl, ks, kr, idx, KV_LORA)
        del tmp
        cfg = SimpleNamespace(scaling=scaling, num_attention_heads=args.heads, kv_lora_rank=KV_LORA,
                              qk_nope_head_dim=NOPE_HEAD, qk_rope_head_dim=ROPE, v_head_dim=KV_LORA, context_len=ctx_len)
        self.model_runner = SimpleNamespace(model_config=cfg, server_args=SimpleNamespace(speculative_algorithm=None),
                                            kv_cache_dtype=torch.float8_e4m3fn, kv_cache_quant_method="per_token_head",
                                            kv_allocator=self.alloc, dtype=torch.bfloat16, device=dev)
        fb.set_attention_tp_size_getter(lambda: tp)    # what dp_attention.get_attention_tp_size() returns in the reference
        try:
            self.backend = fb.FlashMLABackend(self.model_runner)
        finally:
            fb.set_attention_tp_size_getter(None)
        assert self.backend.num_q_heads == h
        self.backend.init_cuda_graph_state(max_bs=b)
        self.attn_layers = [SimpleNamespace(layer_id=l, tp_q_head_num=h, head_dim=KV_LORA + ROPE, v_head_dim=KV_LORA,
                                            scaling=scaling) for l in range(L)]

        # ---- step inputs: host copies (e2e) and the static device tensors the graph reads ----
        last = seq - 1
        self.hidden_host = torch.randn(b, HIDDEN, dtype=torch.bfloat16).pin_memory()
        self.seq_lens_host = torch.full((b,), seq, dtype=torch.int32).pin_memory()
        self.req_pool_host = torch.arange(b, dtype=torch.int64).pin_memory()
        self.positions_host = torch.full((b,), last, dtype=torch.int64).pin_memory()
        loc = self.page_table_host[:, last // PAGE].to(torch.int64) * PAGE + last % PAGE
        self.out_cache_loc_host = loc.to(torch.int32).pin_memory()
        self.result_host = torch.empty(b, HIDDEN, dtype=torch.bfloat16).pin_memory()
        self.hidden_in = self.hidden_host.to(dev)
        self.seq_lens = self.seq_lens_host.to(dev)
        self.req_pool_indices = self.req_pool_host.to(dev)
        self.positions = self.positions_host.to(dev)
        self.out_cache_loc = self.out_cache_loc_host.to(dev)
        self.fbatch = SimpleNamespace(batch_size=b, seq_lens=self.seq_lens, req_pool_indices=self.req_pool_indices,
                                      out_cache_loc=self.out_cache_loc, token_to_kv_pool=self.pool,
                                      forward_mode=fb.ForwardMode.DECODE, spec_info=None)
        inv_freq = 1.0 / (10000.0 ** (torch.arange(0, ROPE, 2, dtype=torch.float32) / ROPE))
        freqs = torch.arange(seq, dtype=torch.float32)[:, None] * inv_freq[None, :]
        self.cos_sin_cache = torch.cat([freqs.cos(), freqs.sin()], dim=-1).contiguous().to(dev)

        # ---- weights (random init, the reference's --load-format dummy) ----
        self.method = fl.Fp8LinearMethod(fl.Fp8Config(True, "dynamic", None, [128, 128]), tp_size=tp)

        def linear(n, k, n_full, k_full):
            layer = torch.nn.Module()
            self.method.create_weights(layer, k, [n], k_full, n_full, torch.bfloat16, device=dev)
            w, s = fl.dummy_block_fp8_weight(n, k, dev, g)
            layer.weight.data.copy_(w)
            layer.weight_scale_inv.data.copy_(s)
            self.method.process_weights_after_loading(layer)
            return layer

        qk = NOPE_HEAD + ROPE
        self.lin_qkv_a = [linear(Q_LORA + KV_LORA + ROPE, HIDDEN, Q_LORA + KV_LORA + ROPE, HIDDEN) for _ in range(L)]  # replicated
        self.lin_q_b = [linear(h * qk, Q_LORA, args.heads * qk, Q_LORA) for _ in range(L)]             # column-parallel
        self.lin_o = [linear(HIDDEN, h * V_HEAD, HIDDEN, args.heads * V_HEAD) for _ in range(L)]      # row-parallel
        self.norms = []
        for _ in range(L):
            n = ln.FusedRMSNorm(ln.RMSNorm(Q_LORA, 1e-6, dev), ln.RMSNorm(KV_LORA, 1e-6, dev))
            n.q_a_norm.weight.data.copy_((torch.rand(Q_LORA, device=dev, generator=g) + 0.5).to(torch.bfloat16))
            n.kv_a_norm.weight.data.copy_((torch.rand(KV_LORA, device=dev, generator=g) + 0.5).to(torch.bfloat16))
            self.norms.append(n)
        bf = lambda *shape, std: (torch.randn(*shape, device=dev, generator=g) * std).to(torch.bfloat16)
        self.w_kc = [bf(h, NOPE_HEAD, KV_LORA, std=NOPE_HEAD ** -0.5) for _ in range(L)]   # deepseek_v2.py:1629-1633
        self.w_vc = [bf(h, KV_LORA, V_HEAD, std=KV_LORA ** -0.5) for _ in range(L)]
        self.q_buf = torch.empty(b, h, KV_LORA + ROPE, dtype=torch.bfloat16, device=dev)  # "Q" of forward_absorb (:829-833)
        self.launches_per_step = 0
        self.library_launches_per_step = 0
        self.backend.init_forward_metadata_capture_cuda_graph(b, b, self.req_pool_indices, self.seq_lens,
                                                              fb.ForwardMode.DECODE, None)
        # direct-kernel state for the roofline microbenchmark of the dominant kernel
        self.attn_out = torch.empty(b, 1, h, KV_LORA, dtype=torch.bfloat16, device=dev)
        self.lse = torch.empty(b, h, 1, dtype=torch.float32, device=dev)
        from sglang_fluentllm_b200 import _lib
        L_ = _lib.lib()
        parts = L_.b200_mla_num_sm_parts(h, 1, L_.b200_sm_count())
        self.workspace = torch.empty(L_.b200_mla_workspace_bytes(b, 1, h, parts), dtype=torch.uint8, device=dev)

    def host_to_device(self):
        """The step's inputs from pinned host memory (e2e): hidden states, sequence lengths, request slots, positions,
        the new tokens' cache slots and the scheduler's page-table rows."""
        nb = True
        self.hidden_in.copy_(self.hidden_host, non_blocking=nb)
        self.seq_lens.copy_(self.seq_lens_host, non_blocking=nb)
        self.req_pool_indices.copy_(self.req_pool_host, non_blocking=nb)
        self.positions.copy_(self.positions_host, non_blocking=nb)
        self.out_cache_loc.copy_(self.out_cache_loc_host, non_blocking=nb)
        self.alloc.req_to_page[: self.args.batch, : self.page_table_host.shape[1]].copy_(self.page_table_host, non_blocking=nb)

    @property
    def h2d_bytes(self):
        ts = (self.hidden_host, self.seq_lens_host, self.req_pool_host, self.positions_host, self.out_cache_loc_host,
              self.page_table_host)
        return sum(t.numel() * t.element_size() for t in ts)

    def replay_metadata(self):
        """cuda_graph_runner.py:505-513: outside the graph, before every replay."""
        self.backend.init_forward_metadata_replay_cuda_graph(self.args.batch, self.req_pool_indices, self.seq_lens,
                                                             self.args.batch * self.args.seq, self.fb.ForwardMode.DECODE,
                                                             None, seq_lens_cpu=None)

    def layer(self, l, x):
        torch = self.torch
        b, h = self.args.batch, self.h
        m = self.method
        # fused_qkv_a_proj_with_mqa (deepseek_v2.py:796-799): activation quantiser + block-FP8 GEMM
        qkv = m.apply(self.lin_qkv_a[l], x)
        # fused q_a / kv_a RMSNorm emitting the q_b_proj input block-quantised (:800-807,816-820)
        q8, block_scale = self.norms[l].forward_block_quant(qkv, Q_LORA)
        q = m.apply(self.lin_q_b[l], q8, block_scale=block_scale, output_dtype=torch.bfloat16).view(b, h, NOPE_HEAD + ROPE)
        q_nope, q_pe = q.split([NOPE_HEAD, ROPE], dim=-1)
        # w_kc absorb (:829-840): bf16 bmm straight into the latent columns of Q, as the reference does it
        Q = self.q_buf
        torch.bmm(q_nope.transpose(0, 1), self.w_kc[l], out=Q[..., :KV_LORA].transpose(0, 1))
        # rotary embedding with the fused set_kv_buffer argument (:841-857): rope of q_pe / k_pe, Q quantiser and
        # quantise-and-cache of the new token's K in one launch; returns the quantised Q triple
        latent_cache = qkv[:, Q_LORA:]
        arg = self.rope_mod.create_fused_set_kv_buffer_arg(value=latent_cache[:, :KV_LORA], layer=self.attn_layers[l],
                                                          forward_batch=self.fbatch)
        triple = self.rope_mod.fused_rope_quantize_set_kv(self.positions, Q[..., :KV_LORA], q_pe, latent_cache[:, KV_LORA:],
                                                          self.cos_sin_cache, False, arg)
        # attn_mqa -> AttentionBackend.forward_decode (:872-884), save_kv_cache=False: the prologue cached K already
        attn = self.backend.forward_decode(triple, None, None, self.attn_layers[l], self.fbatch, save_kv_cache=False)
        # w_vc absorb (:886-887)
        attn_bmm = torch.bmm(attn.view(b, h, KV_LORA).transpose(0, 1), self.w_vc[l])            # [h, b, 128]
        # o_proj (:888): the [b, h, 128] view of the bmm output is quantised in place (no flatten copy)
        xq, xs = self.fk.per_token_group_quant_fp8_strided(attn_bmm.transpose(0, 1))
        y = m.apply(self.lin_o[l], xq, block_scale=xs, output_dtype=torch.bfloat16)
        if self.tp > 1:
            self.tpmod.all_reduce_sum(y)  # the path's single exchange: row-parallel partial sums over NVLink
        # launches: act-quant, GEMM, norm+quant, GEMM, bmm, prologue, decode (merges cut requests itself), bmm, act-quant,
        # GEMM + split-K reduce
        self.launches_per_step += 9     # this repo's kernels
        self.library_launches_per_step += 2   # the two cuBLAS bmm's
        return y

    def step(self, x):
        self.launches_per_step = 0
        self.library_launches_per_step = 0
        for l in range(self.args.layers):
            x = self.layer(l, x)
        return x

    def attn_only(self, l):
        """The dominant kernel alone, through the kernel-module API the backend calls."""
        n_pages = self.n_pages
        if not hasattr(self, "_qq"):
            q_abs = self.torch.randn(self.args.batch, 1, self.h, KV_LORA + ROPE, dtype=self.torch.bfloat16, device=self.dev)
            self._qq = self.fm.quantize_ckv_per_token_head(q_abs, KV_LORA)
        q_nope, q_scale, q_rope = self._qq
        kl, ks, kr = self.pool.get_key_buffer(l)
        md = self.backend.forward_metadata
        self.fm.flash_mla_ckv_fp8_per_token(
            q_nope, q_rope, kl.view(-1, PAGE, 1, KV_LORA), kr.view(-1, PAGE, 1, ROPE), q_scale, ks.view(-1, PAGE, 1, 1),
            md.block_table[: self.args.batch], self.seq_lens, KV_LORA, md.flashmla_metadata, md.num_splits,
            softmax_scale=self.attn_layers[l].scaling, out=self.attn_out, lse=self.lse, workspace=self.workspace)


def workload_config(args, world, parallelism, heads_per_gpu, cuda_graph=True):
    """`config` of the JSON line -- the same dictionary for the B200 arm and for `--impl reference`."""
    n_repl = world if parallelism == "dp" else 1
    bytes_, _ = mla_bytes_flops(args.batch, args.seq, heads_per_gpu)
    return {"workload": "DeepSeek-V3 attention block decode (forward_absorb), %d layers: qkv_a / q_b / o_proj "
                        "block-FP8 GEMMs, q_a / kv_a RMSNorm, w_kc / w_vc absorb, rope + quantise + cache prologue, "
                        "paged MLA decode over per-token-FP8 KV" % args.layers,
            "global_batch": args.batch * n_repl, "seq_len": args.seq, "heads_per_gpu": heads_per_gpu,
            "parallelism": "%s%d" % (parallelism, world), "page_size": PAGE, "cuda_graph": cuda_graph,
            "api": "FlashMLABackend.forward_decode (CUDA-graph capture/replay protocol) + Fp8LinearMethod.apply + "
                   "FusedRMSNorm + rotary prologue with fused set_kv_buffer + w_kc / w_vc bmm (torch.bmm, as the reference)",
            "l2": "inputs larger than L2: %.1f GB of distinct KV + weights streamed per step" % (args.layers * bytes_ / 1e9)}


def mla_bytes_flops(b, seq, h, s_q=1):
    """reference formulas tests/test_flash_mla_ckv_fp8_per_token.py:123-132."""
    S, M = b * seq, h * s_q
    return S * 644 + b * M * 644 + b * M * 1024, 2 * S * M * (576 + 512)


def log(msg):
    if os.environ.get("BENCH_VERBOSE"):
        print("[bench rank %s %.1fs] %s" % (os.environ.get("RANK", "0"), time.time() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.time()


def run_b200(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    log("process group up")
    if args.parallelism is None:
        args.parallelism = "dp" if world > 1 else "tp"
    hp = HotPath(args, rank, world, dev)
    log("state built")
    torch.cuda.synchronize()  # the state was built on the default stream; the step runs on its own stream
    stream = torch.cuda.Stream(device=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.cuda.stream(stream):
        # eager warm-up (also triggers one-time cudaFuncSetAttribute calls) then graph capture
        for _ in range(2):
            hp.step(hp.hidden_in)
        stream.synchronize()
        log("eager warm-up done")
        if args.profile_step:
            # launch list of exactly one step: run under
            #   ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...
            torch.cuda.profiler.start()
            hp.step(hp.hidden_in)
            stream.synchronize()
            torch.cuda.profiler.stop()
            print(json.dumps({"profile_step": True, "gpu_launches": hp.launches_per_step}))
            return
        graph = None
        if not args.no_graph:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=stream):
                hp.x_out = hp.step(hp.hidden_in)
        launches = hp.launches_per_step
        log("graph captured")

        def run_step():
            if graph is not None:
                graph.replay()
            else:
                hp.x_out = hp.step(hp.hidden_in)

        def run_e2e_step():
            # what the model runner does per decode step around the graph (cuda_graph_runner.py:460-526): inputs from
            # pinned host memory into the static tensors, the backend's replay-metadata refresh (block-table gather +
            # tile-scheduler kernel), replay, result back to the host
            hp.host_to_device()
            hp.replay_metadata()
            run_step()
            hp.result_host.copy_(hp.x_out, non_blocking=True)

        def timed(fn, steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
            stream.synchronize()
            barrier()
            return hp.tpmod.max_over_ranks_ms(e0.elapsed_time(e1), dev)

        for _ in range(max(args.warmup, 3)):
            run_step()
        log("warm-up replays done")
        if args.timeline:
            write_timeline(args.timeline, run_step, stream, args.layers)
            return
        sampler = ClockSampler(local_rank)
        sampler.start()
        ms = timed(run_step, args.steps)
        clocks = sampler.stop()
        log("timed region done")
        for _ in range(2):
            run_e2e_step()
        ms_e2e = timed(run_e2e_step, args.steps)
        log("e2e region done")

        # dominant kernel: MLA decode (+combine), timed alone per layer (each layer's KV is distinct and the set of
        # layers is far larger than L2, so every launch is L2-cold)
        for l in range(min(3, args.layers)):
            hp.attn_only(l)
        stream.synchronize()
        n_attn = min(args.layers, 32)
        attn_graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(attn_graph, stream=stream):
            for l in range(n_attn):
                hp.attn_only(l)
        attn_graph.replay()
        stream.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(3):
            attn_graph.replay()
        e1.record(stream)
        stream.synchronize()
        attn_ms = e0.elapsed_time(e1) / (3 * n_attn)

    ms_step = ms / args.steps
    n_repl = world if args.parallelism == "dp" else 1
    tokens_per_s = args.batch * n_repl / (ms_step * 1e-3)
    e2e_tokens_per_s = args.batch * n_repl / (ms_e2e / args.steps * 1e-3)
    bytes_, flops = mla_bytes_flops(args.batch, args.seq, hp.h)
    achieved = bytes_ / (attn_ms * 1e-3) / 1e9
    h2d = hp.h2d_bytes
    d2h = hp.result_host.numel() * 2
    traffic = traffic_source = None
    try:  # DRAM traffic of the dominant kernel from the committed ncu capture of the same shape (profiles/)
        key = "mla_decode_pair_kernel b=%d seq=%d h=%d s_q=1" % (args.batch, args.seq, hp.h)
        entry = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[key]
        traffic, traffic_source = entry["traffic_bytes"], entry["source"]
    except Exception:
        traffic, traffic_source = None, None
    line = {
        "metric": METRIC, "value": round(tokens_per_s, 2),
        "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_step, 4), "higher_is_better": True,
        "scaling": "strong" if (args.parallelism == "tp" and world > 1) else "weak", "vs_baseline": None,
        "dtype": "fp8_e4m3",
        "data": "synthetic (random-init weights, N(0,1) KV quantised by the path's own quantiser, constant seq len)",
        "config": workload_config(args, world, args.parallelism, hp.h, graph is not None),
        "e2e": {"value": round(e2e_tokens_per_s, 2), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches * args.steps,  # this repo's kernels only
        "library_launches": hp.library_launches_per_step * args.steps,  # cuBLAS bf16 bmm (w_kc / w_vc absorb, as the reference)
        "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(achieved / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_source,
                     "peak_source": peak_src,
                     "kernel": "mla_decode_pair_kernel (merges cut requests itself; + combine launch where the backend needs one), per layer, L2-cold, model softmax scale", "kernel_ms": round(attn_ms, 4),
                     "algorithmic_bytes": bytes_, "tflops": round(flops / (attn_ms * 1e-3) / 1e12, 1),
                     # the second roofline (informational): the page-wise P arithmetic issues P.V twice (hi + lo weights,
                     # DESIGN 3.1), so the tensor pipe executes 2 S M (576 + 2 * 512) flops; peaks = torch._scaled_mm
                     # 8192^3 on this pool (profiles/r2h_gemm_compare.txt: burst / sustained under the power cap)
                     "tensor": {"executed_tflops": round(flops * 1600 / 1088 / (attn_ms * 1e-3) / 1e12, 1),
                                "fp8_peak_burst": 3107.0, "fp8_peak_sustained": 2612.0,
                                "frac_of_sustained": round(flops * 1600 / 1088 / (attn_ms * 1e-3) / 1e12 / 2612.0, 4)}},
    }
    tp_strong = None
    if world > 1 and args.parallelism == "dp":
        # the same step under strict head tensor parallelism (global batch fixed, one NCCL all-reduce per layer)
        graph = attn_graph = hp = None
        torch.cuda.empty_cache()
        import copy
        targs = copy.copy(args)
        targs.parallelism = "tp"
        hp_tp = HotPath(targs, rank, world, dev)
        torch.cuda.synchronize()
        # the all-reduce: this repo's two-shot kernel over NVLink peer memory (csrc/p2p_comm.cu); B200_TP_NCCL=1 keeps
        # torch.distributed's NCCL call for A/B
        use_p2p = not os.environ.get("B200_TP_NCCL")
        if use_p2p:
            try:
                hp_tp.tpmod.enable_p2p_all_reduce(None, dev, args.batch * HIDDEN * 2)
                ok = torch.ones(1, device=dev)
            except Exception as e:  # no peer access / IPC on this box: say so and time the NCCL form
                log("peer-memory all-reduce unavailable (%s: %s): tp_strong uses NCCL" % (type(e).__name__, e))
                ok = torch.zeros(1, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)  # every rank or none
            if float(ok.item()) == 0.0:
                use_p2p = False
                hp_tp.tpmod._p2p = None
        with torch.cuda.stream(stream):
            for _ in range(2):
                hp_tp.step(hp_tp.hidden_in)
            stream.synchronize()
            g_tp = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_tp, stream=stream):
                hp_tp.x_out = hp_tp.step(hp_tp.hidden_in)
            for _ in range(3):
                g_tp.replay()
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(args.steps):
                g_tp.replay()
            e1.record(stream)
            stream.synchronize()
            barrier()
            ms_tp = hp_tp.tpmod.max_over_ranks_ms(e0.elapsed_time(e1), dev) / args.steps
        tp_strong = {"value": round(args.batch / (ms_tp * 1e-3), 2), "unit": "tokens/s", "ms_per_step": round(ms_tp, 4),
                     "global_batch": args.batch, "heads_per_gpu": hp_tp.h, "scaling": "strong",
                     "collective": ("1 all-reduce [b, 7168] bf16 per layer on the row-parallel o_proj output: "
                                    + (("this repo's kernel over NVLink (csrc/p2p_comm.cu): "
                                        + ("in-switch reduction (multimem)" if getattr(hp_tp.tpmod._p2p, "_nvls", None)
                                           and world > 2 else "peer loads")) if use_p2p else "NCCL")),
                     "nvlink_bytes_per_step_per_gpu": int(2 * (world - 1) / world * args.batch * HIDDEN * 2 * args.layers)}
        g_tp = hp_tp = None
        log("tp_strong done")
    if tp_strong is not None:
        line["tp_strong"] = tp_strong
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args, budget_s=20.0)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Tearing the NCCL communicator down while captured graphs still reference it hangs (seen with torch 2.11 /
        # NCCL 2.28 at N = 2): drop the graphs, drain the device, meet the other ranks, then leave without the
        # communicator teardown.
        graph = attn_graph = None
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ----------------------------------------------------------------------------------------------------------------
def write_timeline(path, run_step, stream, layers):
    """Per-kernel durations and inter-kernel gaps of one replayed step (CUPTI activity records)."""
    import torch
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run_step()
        stream.synchronize()
    evs = []
    for e in prof.events():
        if getattr(e, "device_type", None) is not None and "CUDA" in str(e.device_type):
            evs.append((e.time_range.start, e.time_range.end, e.name))
    evs.sort()
    per = {}
    busy = 0.0
    gaps = []
    for i, (t0, t1, name) in enumerate(evs):
        d = per.setdefault(name, [0, 0.0, 0.0])  # count, total duration, total gap in front of it
        d[0] += 1
        d[1] += t1 - t0
        busy += t1 - t0
        if i:
            g = t0 - evs[i - 1][1]
            d[2] += g
            gaps.append(g)
    span = evs[-1][1] - evs[0][0] if evs else 0.0
    with open(path, "w") as f:
        f.write("one replayed step: %d device activities, span %.1f us, busy %.1f us, idle between kernels %.1f us (%.1f %%)\n"
                % (len(evs), span, busy, span - busy, 100.0 * (span - busy) / max(span, 1e-9)))
        f.write("%6s %10s %10s %10s  %s\n" % ("count", "avg us", "total us", "gap before", "kernel"))
        for name, (n, tot, gap) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("%6d %10.2f %10.1f %10.2f  %s\n" % (n, tot / n, tot, gap / n, name[:110]))
        per_layer = len(evs) // max(layers, 1)
        if per_layer and len(evs) >= 3 * per_layer:
            f.write("layer 2 in launch order (start relative to the layer's first kernel, duration, gap before):\n")
            base = evs[2 * per_layer][0]
            for i in range(2 * per_layer, 3 * per_layer):
                t0, t1, name = evs[i]
                f.write("  %9.2f %8.2f %7.2f  %s\n" % (t0 - base, t1 - t0, t0 - evs[i - 1][1], name[:100]))


def cpu_baseline(args, budget_s=20.0):
    """The reference's CPU path for this step (torch-native SDPA decode loop over a bf16 pool + native block-FP8
    matmul), ported in oracle/native_cpu.py, on a bounded sample: `sample_b` requests of one layer."""
    import torch

    from oracle import native_cpu, ref_import
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # the reference's own functions where /root/reference exists (the build container), their port elsewhere (GPU box)
    ref = ref_import.reference_cpu_path()
    decode, matmul, quant = ref if ref is not None else (None, None, None)
    sample_b = min(32, args.batch)  # at least a quarter of the batch; doubled while the budget allows
    t_attn = native_cpu.time_decode_attention(sample_b, args.seq, args.heads, repeats=1, decode=decode)
    while t_attn * 4 < budget_s / 2 and sample_b < args.batch:
        sample_b = min(2 * sample_b, args.batch)
        t_attn = native_cpu.time_decode_attention(sample_b, args.seq, args.heads, repeats=1, decode=decode)
    t_gemm = native_cpu.time_block_fp8_linears(min(args.batch, 32), args.heads, repeats=1, matmul=matmul,
                                               quant=quant) * (args.batch / min(args.batch, 32))
    per_layer = t_attn * (args.batch / sample_b) + t_gemm
    value = args.batch / (per_layer * args.layers)
    return {"value": round(value, 3), "unit": "tokens/s", "cores": cores, "kind": "reference" if ref is not None else "port",
            "sample": "1 layer: torch-native SDPA decode (%s) for %d of %d requests (seq %d, %d heads, bf16 pool) + native "
                      "block-FP8 matmuls for %d rows; scaled to %d requests x %d layers" % (
                          "TorchNativeAttnBackend._run_sdpa_forward_decode, the reference's own source"
                          if ref is not None else "port of TorchNativeAttnBackend._run_sdpa_forward_decode",
                          sample_b, args.batch, args.seq, args.heads, min(args.batch, 32), args.batch, args.layers)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for _ in range(max(1, min(args.steps, 3))):
        base = cpu_baseline(args, budget_s=15.0)
        vals.append(base["value"])
    v = sorted(vals)[len(vals) // 2]
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(args.batch / v * 1e3, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16/f32 (CPU)",
            "data": "synthetic (same shapes as the B200 arm; the CPU path runs on a bf16 pool: torch-native attention "
                    "cannot read the per-token-FP8 layout)",
            "config": workload_config(args, max(1, args.gpus), "dp" if args.gpus > 1 else "tp",
                                      args.heads, True),
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
