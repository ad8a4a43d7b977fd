"""The fused RMSNorm of the MLA attention block, mirroring ``FusedRMSNorm``
(python/sglang/srt/layers/layernorm.py:249-330): q_a and kv_a of the fused qkv_a projection output are normalised in
one launch (``forward``, called at models/deepseek_v2.py:816-820), optionally emitting the q_b_proj input already
block-quantised (the ``fuse_block_quant_fp8`` epilogue of ``forward_with_allgather_fusion``: the returned scales are the
``block_scale`` the next linear takes, deepseek_v2.py:796-807) -- the activation quantiser launch in front of q_b_proj
disappears."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib
from .fp8_kernel import create_per_token_group_quant_fp8_output_scale


class RMSNorm(torch.nn.Module):
    """The weight / epsilon holder FusedRMSNorm takes two of (layernorm.py: RMSNorm)."""

    def __init__(self, hidden_size: int, eps: float = 1e-6, device=None):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(hidden_size, dtype=torch.bfloat16, device=device), requires_grad=False)
        self.variance_epsilon = eps


class FusedRMSNorm(torch.nn.Module):
    def __init__(self, q_a_norm: RMSNorm, kv_a_norm: RMSNorm):
        super().__init__()
        self.q_a_norm, self.kv_a_norm = q_a_norm, kv_a_norm

    @property
    def weight_q_a(self):
        return self.q_a_norm.weight

    @property
    def weight_kv_a(self):
        return self.kv_a_norm.weight

    def _call(self, qkv: torch.Tensor, q_dim: int, q_out, q8, q_scales):
        if not qkv.is_cuda:
            raise RuntimeError("FusedRMSNorm: CUDA tensor required (no CPU fallback)")
        if qkv.dtype != torch.bfloat16 or qkv.dim() != 2 or qkv.stride(1) != 1:
            raise RuntimeError("FusedRMSNorm: bf16 [tokens, q_lora + kv_lora (+ rope)] with contiguous rows expected")
        m = qkv.shape[0]
        with torch.cuda.device(qkv.device):
            _lib.check(_lib.lib().b200_qk_rmsnorm_quant(
                qkv.data_ptr(), m, qkv.stride(0) if m > 1 else qkv.shape[1], q_dim, self.weight_kv_a.numel(),
                self.weight_q_a.data_ptr(), self.weight_kv_a.data_ptr(), float(self.q_a_norm.variance_epsilon),
                q_out.data_ptr() if q_out is not None else None, q_out.stride(0) if q_out is not None and m > 1 else q_dim,
                q8.data_ptr() if q8 is not None else None, q_scales.data_ptr() if q_scales is not None else None,
                (q_scales.stride(1) if q_scales.shape[1] > 1 else (m + 3) // 4 * 4) if q_scales is not None else 0,
                1e-10, _lib.current_stream()), "qk_rmsnorm_quant")

    def forward(self, input_q_a: torch.Tensor, input_kv_a: torch.Tensor, output_q_a: Optional[torch.Tensor] = None,
                output_kv_a: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        """layernorm.py:276-303.  ``input_q_a`` / ``input_kv_a`` must be the adjacent column slices of one qkv_a output
        row (what ``qkv.split`` gives the reference, deepseek_v2.py:812-815); kv_a is normalised in place."""
        q_dim = input_q_a.shape[-1]
        if output_kv_a is not None and output_kv_a.data_ptr() != input_kv_a.data_ptr():
            raise RuntimeError("FusedRMSNorm: kv_a is normalised in place (the reference's call passes no output_kv_a)")
        if input_kv_a.data_ptr() != input_q_a.data_ptr() + q_dim * 2 or input_kv_a.stride(0) != input_q_a.stride(0):
            raise RuntimeError("FusedRMSNorm: q_a and kv_a must be adjacent column slices of the fused qkv_a output")
        qkv = input_q_a.as_strided((input_q_a.shape[0], q_dim + input_kv_a.shape[-1]), (input_q_a.stride(0), 1))
        out = output_q_a if output_q_a is not None else input_q_a
        self._call(qkv, q_dim, out, None, None)
        return out, input_kv_a

    def forward_block_quant(self, qkv: torch.Tensor, q_dim: int, output_q_a: Optional[torch.Tensor] = None):
        """Both norms + the 1x128 quantisation of the normalised q_a (fuse_block_quant_fp8, layernorm.py:305-330):
        returns (q8 e4m3 [tokens, q_dim], block_scale f32 [tokens, q_dim / 128] column-major TMA-aligned) -- the pair
        ``Fp8LinearMethod.apply(layer, q8, block_scale=...)`` takes."""
        m = qkv.shape[0]
        q8 = torch.empty((m, q_dim), dtype=torch.float8_e4m3fn, device=qkv.device)
        scales = create_per_token_group_quant_fp8_output_scale((m, q_dim), qkv.device, 128, True, True)
        self._call(qkv, q_dim, output_q_a, q8, scales)
        return q8, scales
