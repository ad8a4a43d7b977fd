"""GPU parity tests of the fused decode prologue (rotary embedding + quantize_ckv_per_token_head
+ quantize_and_cache_k in one launch, ``b200_mla_rope_quantize_fused``), through the C ABI.

Bar: BIT-EXACT (e4m3 codes, fp32 scale bits, bf16 bits, pool contents) against
  (1) the oracle (oracle/rope.py: the reference's DeepseekScalingRotaryEmbedding.forward_native arithmetic, pinned by
      tests/golden/rope.npz, composed with the quantiser restatements), and
  (2) the unfused sequence on the GPU: torch rope in fp32 -> bf16 -> the a-6 / a-7 kernels;
and the backend produces the same attention output from the fused triple as from the unfused call."""
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import rope as orope

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def mods():
    from sglang_fluentllm_b200 import _lib, flash_mla_fp8, rope
    _lib.lib()
    return rope, flash_mla_fp8


def _cos_sin_cache(max_pos, seed):
    """A cache with the layout of rotary_embedding.py:791-802 (cos | sin, f32); yarn-like angles."""
    g = torch.Generator().manual_seed(seed)
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, 64, 2, dtype=torch.float32) / 64))
    freqs = torch.arange(max_pos, dtype=torch.float32)[:, None] * inv_freq[None, :]
    mscale = 1.0 + 0.1 * torch.rand(1, generator=g).item()
    return (torch.cat([freqs.cos(), freqs.sin()], dim=-1) * mscale).contiguous()


def _torch_rope(x, positions, cache, neox):
    """x bf16 [t, h, 64] -> bf16, each fp32 operation its own kernel (no contraction): rotary_embedding.py:804-846."""
    cs = cache[positions]
    cos, sin = cs[:, None, :32], cs[:, None, 32:]
    xf = x.float()
    if neox:
        cos, sin = torch.cat([cos, cos], -1), torch.cat([sin, sin], -1)
        rot = torch.cat([-xf[..., 32:], xf[..., :32]], -1)
    else:
        cos, sin = cos.repeat_interleave(2, -1), sin.repeat_interleave(2, -1)
        rot = torch.stack([-xf[..., 1::2], xf[..., 0::2]], -1).flatten(-2)
    a = xf * cos
    b = rot * sin
    return (a + b).to(torch.bfloat16)


def _inputs(tokens, heads, seed, strided):
    torch.manual_seed(seed)
    if strided:
        # the views a model hands over: Q[..., :512] of the [t, h, 576] absorb buffer, q.split(...)[1] of [t, h, 192],
        # the kv columns of the fused qkv_a projection output [t, 1536 + 576]
        q_buf = (torch.randn(tokens, heads, 576, device=DEV) * 2).to(torch.bfloat16)
        q_full = (torch.randn(tokens, heads, 192, device=DEV) * 2).to(torch.bfloat16)
        qkv_a = (torch.randn(tokens, 2112, device=DEV) * 2).to(torch.bfloat16)
        return q_buf[..., :512], q_full[..., 128:], qkv_a[:, 1536:]
    q_nope = (torch.randn(tokens, heads, 512, device=DEV) * 2).to(torch.bfloat16)
    q_pe = (torch.randn(tokens, heads, 64, device=DEV) * 2).to(torch.bfloat16)
    kv = (torch.randn(tokens, 576, device=DEV) * 2).to(torch.bfloat16)
    return q_nope, q_pe, kv


def _pool(slots):
    return (torch.zeros(slots, 512, dtype=torch.uint8, device=DEV), torch.zeros(slots, dtype=torch.float32, device=DEV),
            torch.zeros(slots, 64, dtype=torch.bfloat16, device=DEV))


@pytest.mark.parametrize("neox", [False, True])
@pytest.mark.parametrize("tokens,heads,strided", [(1, 1, False), (5, 16, True), (128, 128, True), (37, 3, False)])
def test_fused_prologue_bit_exact(mods, neox, tokens, heads, strided):
    rope, fm = mods
    max_pos, slots = 8192, 64 * 40
    cache = _cos_sin_cache(max_pos, seed=tokens).to(DEV)
    q_nope, q_pe, kv = _inputs(tokens, heads, seed=heads + tokens, strided=strided)
    q_nope[0, 0] = 0  # an all-zero latent row: scale = 1e-12, the rope dims are stored as rope * 1e12
    positions = torch.randint(0, max_pos, (tokens,), device=DEV)
    positions[0] = 0
    cache_loc = (torch.randperm(slots - 64)[:tokens] + 64).to(torch.int32).to(DEV)
    k_lora, k_scale, k_rope = _pool(slots)
    arg = rope.FusedSetKVBufferArg(value=kv[:, :512], k_buffer=k_rope, v_buffer=k_lora, scale_buffer=k_scale,
                                   cache_loc=cache_loc)
    got_nope, got_scale, got_rope = rope.fused_rope_quantize_set_kv(positions, q_nope, q_pe, kv[:, 512:], cache, neox, arg)
    torch.cuda.synchronize()

    # (1) oracle
    o_lora = np.zeros((slots, 1, 512), np.uint8)
    o_scale = np.zeros((slots, 1, 1), np.float32)
    o_rope = np.zeros((slots, 1, 64), np.float32)
    with np.errstate(over="ignore", invalid="ignore"):
        ref_nope, ref_scale, ref_rope = orope.rope_quantize_fused(
            q_nope.float().cpu().numpy(), q_pe.float().cpu().numpy(), kv.float().cpu().numpy(), positions.cpu().numpy(),
            cache.cpu().numpy(), neox, cache_loc.cpu().numpy(), o_lora, o_scale, o_rope)
    assert np.array_equal(got_nope.cpu().numpy(), ref_nope)
    assert np.array_equal(got_scale.cpu().numpy().view(np.uint32), ref_scale.view(np.uint32))
    assert np.array_equal(got_rope.float().cpu().numpy().view(np.uint32), ref_rope.view(np.uint32))
    assert np.array_equal(k_lora.cpu().numpy(), o_lora[:, 0])
    assert np.array_equal(k_scale.cpu().numpy().view(np.uint32), o_scale[:, 0, 0].view(np.uint32))
    assert np.array_equal(k_rope.float().cpu().numpy().view(np.uint32), o_rope[:, 0].view(np.uint32))

    # (2) the unfused sequence on the GPU
    q = torch.cat([q_nope, _torch_rope(q_pe, positions, cache, neox)], dim=-1).contiguous()
    u_nope, u_scale, u_rope = fm.quantize_ckv_per_token_head(q, 512)
    key = torch.cat([kv[:, :512], _torch_rope(kv[:, None, 512:], positions, cache, neox)[:, 0]], dim=-1).contiguous()
    u_lora, u_ks, u_kr = _pool(slots)
    fm.quantize_and_cache_k(key, u_lora, u_ks, u_kr, cache_loc, 512)
    torch.cuda.synchronize()
    assert torch.equal(got_nope, u_nope) and torch.equal(got_scale.view(torch.int32), u_scale.view(torch.int32))
    assert torch.equal(got_rope.view(torch.int16), u_rope.view(torch.int16))
    assert torch.equal(k_lora, u_lora) and torch.equal(k_scale.view(torch.int32), u_ks.view(torch.int32))
    assert torch.equal(k_rope.view(torch.int16), u_kr.view(torch.int16))


def test_fused_prologue_q_only_and_errors(mods):
    rope, _ = mods
    cache = _cos_sin_cache(256, seed=3).to(DEV)
    q_nope, q_pe, kv = _inputs(4, 8, seed=9, strided=False)
    positions = torch.tensor([0, 5, 255, 17], device=DEV, dtype=torch.int32)   # int32 positions are widened (rope.py:1237)
    n, s, r = rope.fused_rope_quantize_set_kv(positions, q_nope, q_pe, kv[:, 512:], cache, False, None)
    torch.cuda.synchronize()
    ref_n, ref_s, ref_r = orope.rope_quantize_fused(q_nope.float().cpu().numpy(), q_pe.float().cpu().numpy(), None,
                                                    positions.cpu().numpy(), cache.cpu().numpy(), False)
    assert np.array_equal(n.cpu().numpy(), ref_n) and np.array_equal(s.cpu().numpy(), ref_s)
    assert np.array_equal(r.float().cpu().numpy(), ref_r)
    with pytest.raises(ValueError):
        rope.fused_rope_quantize_set_kv(positions, q_nope, q_pe, kv[:, 512:], cache.double(), False, None)
    with pytest.raises(RuntimeError):
        rope.fused_rope_quantize_set_kv(positions.cpu(), q_nope, q_pe, kv[:, 512:], cache, False, None)
    lora, scale, kr = _pool(128)
    bad = rope.FusedSetKVBufferArg(value=kv[:, :512].contiguous(), k_buffer=kr, v_buffer=lora, scale_buffer=scale,
                                   cache_loc=torch.arange(4, device=DEV))
    with pytest.raises(RuntimeError):   # value and k_pe are not views of the same 576-wide rows
        rope.fused_rope_quantize_set_kv(positions, q_nope, q_pe, kv[:, 512:], cache, False, bad)


def test_backend_decode_from_fused_triple(mods):
    """forward_decode(q = fused triple) == forward_decode(q, k, v) with the unfused prologue, bit for bit."""
    rope, _ = mods
    from test_gpu_backend import _fill_context, _runner
    heads, ctx = 16, [130, 64, 700, 1]
    bs = len(ctx)
    cache = _cos_sin_cache(4096, seed=5).to(DEV)
    q_nope, q_pe, kv = _inputs(bs, heads, seed=21, strided=True)
    positions = torch.tensor(ctx, device=DEV)
    outs = []
    for fused in (False, True):
        fb, mr, alloc, pool, layer = _runner(heads)
        _fill_context(alloc, pool, layer, ctx)
        backend = fb.FlashMLABackend(mr)
        new_loc = torch.cat([alloc.alloc(i, 1, n) for i, n in enumerate(ctx)])
        fbatch = SimpleNamespace(batch_size=bs, seq_lens=torch.tensor([n + 1 for n in ctx], device=DEV),
                                 req_pool_indices=torch.arange(bs, device=DEV), out_cache_loc=new_loc,
                                 token_to_kv_pool=pool, forward_mode=fb.ForwardMode.DECODE)
        backend.init_forward_metadata(fbatch)
        if fused:
            arg = rope.create_fused_set_kv_buffer_arg(kv[:, :512], layer, fbatch)
            triple = rope.fused_rope_quantize_set_kv(positions, q_nope, q_pe, kv[:, 512:], cache, False, arg)
            outs.append(backend.forward_decode(triple, None, None, layer, fbatch, save_kv_cache=False))
        else:
            q = torch.cat([q_nope, _torch_rope(q_pe, positions, cache, False)], dim=-1).reshape(bs, heads * 576)
            k = torch.cat([kv[:, :512], _torch_rope(kv[:, None, 512:], positions, cache, False)[:, 0]], dim=-1)[:, None]
            outs.append(backend.forward_decode(q, k, k[..., :512], layer, fbatch))
    torch.cuda.synchronize()
    assert outs[0].shape == (bs, heads * 512)
    assert torch.equal(outs[0].view(torch.int16), outs[1].view(torch.int16))
