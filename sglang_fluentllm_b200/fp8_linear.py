"""Operator surface of the block-scaled W8A8 FP8 linear: same names, arguments and
error behaviour as the reference, with the B200 kernels behind them.

  * ``Fp8Config``            python/sglang/srt/layers/quantization/fp8.py:46-109
  * ``Fp8LinearMethod``      python/sglang/srt/layers/dense/layouts/fp8.py:71-355 (block-quant branch: create_weights
                             :105-211, process_weights_after_loading :213-232, apply :317-355)
  * ``dense_deep_gemm`` / ``deepgemm_w8a8_block_fp8_linear_with_fallback``
                             python/sglang/srt/layers/dense/gemms/fp8/deep_geem.py:14-102
  * ``dispatch_w8a8_block_fp8_linear``  python/sglang/srt/layers/dense/gemms/fp8/__init__.py:11-17

Only the block-quantised, dynamic-activation path is on the hot path; per-tensor FP8, Marlin and the Triton fall-back
stay the reference's.  There is no fall-back here: a shape or dtype the kernel does not take raises.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional

import torch

from . import deep_gemm
from .fp8_kernel import sglang_per_token_group_quant_fp8

ACTIVATION_SCHEMES = ["static", "dynamic"]  # quantization/fp8.py:40


class Fp8Config:
    """quantization/fp8.py:46-109 (same constructor checks and messages)."""

    def __init__(self, is_checkpoint_fp8_serialized: bool = False, activation_scheme: str = "dynamic",
                 ignored_layers: Optional[List[str]] = None, weight_block_size: Optional[List[int]] = None) -> None:
        self.is_checkpoint_fp8_serialized = is_checkpoint_fp8_serialized
        if activation_scheme not in ACTIVATION_SCHEMES:
            raise ValueError(f"Unsupported activation scheme {activation_scheme}")
        self.activation_scheme = activation_scheme
        self.ignored_layers = ignored_layers or []
        if weight_block_size is not None:
            if not is_checkpoint_fp8_serialized:
                raise ValueError("The block-wise quantization only supports fp8-serialized checkpoint for now.")
            if len(weight_block_size) != 2:
                raise ValueError("The quantization block size of weight must have 2 dimensions, but got "
                                 f"{len(weight_block_size)} dimensions.")
            if activation_scheme != "dynamic":
                raise ValueError("The block-wise quantization only supports dynamic activation scheme for now, but got "
                                 f"{activation_scheme} activation scheme.")
        self.weight_block_size = weight_block_size

    @classmethod
    def get_name(cls) -> str:
        return "fp8"

    @classmethod
    def get_supported_act_dtypes(cls) -> List[torch.dtype]:
        return [torch.bfloat16, torch.half]

    @classmethod
    def get_min_capability(cls) -> int:
        return 100  # this build is sm_100a only (reference: 80)

    @classmethod
    def get_config_filenames(cls) -> List[str]:
        return []

    @staticmethod
    def get_from_keys(config: Dict[str, Any], keys: List[str]) -> Any:
        for key in keys:
            if key in config:
                return config[key]
        raise ValueError(f"Cannot find any of {keys} in the model's quantization config.")

    @classmethod
    def get_from_keys_or(cls, config: Dict[str, Any], keys: List[str], default: Any) -> Any:
        try:
            return cls.get_from_keys(config, keys)
        except ValueError:
            return default

    @classmethod
    def from_config(cls, config: Dict[str, Any]) -> "Fp8Config":
        quant_method = cls.get_from_keys(config, ["quant_method"])
        return cls(is_checkpoint_fp8_serialized="fp8" in quant_method,
                   activation_scheme=cls.get_from_keys(config, ["activation_scheme"]),
                   ignored_layers=cls.get_from_keys_or(config, ["ignored_layers"], None),
                   weight_block_size=cls.get_from_keys_or(config, ["weight_block_size"], None))

    def get_quant_method(self, layer: torch.nn.Module, prefix: str = ""):
        if any(prefix == name or prefix.startswith(name + ".") for name in self.ignored_layers):
            return None
        return Fp8LinearMethod(self)

    def get_scaled_act_names(self) -> List[str]:
        return []


def dummy_block_fp8_weight(n: int, k: int, device, generator: Optional[torch.Generator] = None, std: Optional[float] = None):
    """Random-init block-FP8 weight for ``--load-format dummy`` style runs (the reference initialises dummy weights at
    random instead of reading a checkpoint; an FP8 checkpoint stores e4m3 codes + a 128x128 ``weight_scale_inv``): e4m3
    codes of N(0, 1) values and one positive scale per block such that the de-quantised weight has standard deviation
    ``std`` (default k ** -0.5).  Returns (weight e4m3 [n, k], weight_scale_inv f32 [ceil(n/128), ceil(k/128)])."""
    std = (k ** -0.5) if std is None else std
    codes = torch.randn(n, k, dtype=torch.float32, device=device, generator=generator).to(torch.float8_e4m3fn)
    jitter = torch.rand(-(-n // 128), -(-k // 128), dtype=torch.float32, device=device, generator=generator)
    return codes, (std * (0.75 + 0.5 * jitter)).contiguous()


def dense_deep_gemm(A: torch.Tensor, B: torch.Tensor, As: torch.Tensor, Bs: torch.Tensor, block_size: List[int],
                    output_dtype: torch.dtype = torch.float16, use_pdl: bool = True) -> torch.Tensor:
    """dense/gemms/fp8/deep_geem.py:14-56: shape checks, output allocation, ``gemm_fp8_fp8_bf16_nt``."""
    assert len(block_size) == 2
    block_n, block_k = block_size[0], block_size[1]
    assert A.shape[-1] == B.shape[-1]
    assert A.shape[:-1] == As.shape[:-1] and A.is_contiguous()
    assert -(-A.shape[-1] // block_k) == As.shape[-1], f"{A.shape=} {block_k=} {As.shape=}"
    assert B.ndim == 2 and B.is_contiguous() and Bs.ndim == 2
    N, K = B.shape
    assert -(-N // block_n) == Bs.shape[0]
    assert -(-K // block_k) == Bs.shape[1]
    C = A.new_empty(A.shape[:-1] + (N,), dtype=output_dtype)
    assert C.dtype == torch.bfloat16, f"C.dtype: {C.dtype} != torch.bfloat16"  # deepgemm only supports bf16
    deep_gemm.gemm_fp8_fp8_bf16_nt((A, As), (B, Bs), C, use_pdl)
    return C


def deepgemm_w8a8_block_fp8_linear_with_fallback(input: torch.Tensor, weight: torch.Tensor, block_size: List[int],
                                                 weight_scale: torch.Tensor, input_scale: Optional[torch.Tensor] = None,
                                                 bias: Optional[torch.Tensor] = None,
                                                 output_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """dense/gemms/fp8/deep_geem.py:59-102.  Same flow (1x128 activation quantisation into column-major TMA-aligned
    scales unless the producer already supplies ``input_scale``, GEMM, bias); where the reference falls back to Triton
    (non-bf16 output, n % 64 or k % 128) this build raises: there is no second code path."""
    if output_dtype is None:
        output_dtype = input.dtype
    if output_dtype != torch.bfloat16:
        raise RuntimeError("w8a8_block_fp8_linear: only bfloat16 output is implemented on the B200 path "
                           f"(got {output_dtype}); no fallback")
    if not (weight.shape[0] % 64 == 0 and weight.shape[1] % 128 == 0):
        raise RuntimeError("w8a8_block_fp8_linear: weight shape %s needs n %% 64 == 0 and k %% 128 == 0; no fallback"
                           % (tuple(weight.shape),))
    input_2d = input.view(-1, input.shape[-1])
    output_shape = [*input.shape[:-1], weight.shape[0]]
    if input_scale is None:
        q_input, x_scale = sglang_per_token_group_quant_fp8(input_2d, block_size[1], column_major_scales=True,
                                                            scale_tma_aligned=True)
    else:
        q_input, x_scale = input_2d, input_scale
    output = dense_deep_gemm(q_input, weight, x_scale, weight_scale, block_size, output_dtype=output_dtype)
    if bias is not None:
        output += bias
    return output.to(dtype=output_dtype).view(*output_shape)


def dispatch_w8a8_block_fp8_linear() -> Callable:
    """dense/gemms/fp8/__init__.py:11-17: one backend here."""
    return deepgemm_w8a8_block_fp8_linear_with_fallback


def check_block_alignment(input_size: int, input_size_per_partition: int, output_size: int,
                          output_partition_sizes: List[int], weight_block_size: List[int], tp_size: int) -> None:
    """The tensor-parallel block-alignment rules of Fp8LinearMethod.create_weights (dense/layouts/fp8.py:118-142)."""
    block_n, block_k = weight_block_size
    output_size_per_partition = sum(output_partition_sizes)
    if tp_size > 1 and input_size // input_size_per_partition == tp_size:  # row parallel
        if input_size_per_partition % block_k != 0:
            raise ValueError(f"Weight input_size_per_partition = {input_size_per_partition} is not divisible by "
                             f"weight quantization block_k = {block_k}.")
    if (tp_size > 1 and output_size // output_size_per_partition == tp_size) or len(output_partition_sizes) > 1:
        for output_partition_size in output_partition_sizes:  # column parallel or merged weights
            if output_partition_size % block_n != 0:
                raise ValueError(f"Weight output_partition_size = {output_partition_size} is not divisible by "
                                 f"weight quantization block_n = {block_n}.")


class Fp8LinearMethod:
    """Block-quant branch of dense/layouts/fp8.py:71-355."""

    def __init__(self, quant_config: Fp8Config, tp_size: int = 1):
        self.quant_config = quant_config
        self.block_quant = self.quant_config.weight_block_size is not None
        if not self.block_quant:
            raise NotImplementedError("only block-quantised FP8 (weight_block_size) is on the B200 hot path")
        self.tp_size = tp_size
        self.w8a8_block_fp8_linear = dispatch_w8a8_block_fp8_linear()

    def create_weights(self, layer: torch.nn.Module, input_size_per_partition: int, output_partition_sizes: List[int],
                       input_size: int, output_size: int, params_dtype: torch.dtype, **extra_weight_attrs) -> None:
        block_n, block_k = self.quant_config.weight_block_size
        check_block_alignment(input_size, input_size_per_partition, output_size, output_partition_sizes,
                              self.quant_config.weight_block_size, self.tp_size)
        output_size_per_partition = sum(output_partition_sizes)
        layer.logical_widths = output_partition_sizes
        layer.input_size_per_partition = input_size_per_partition
        layer.output_size_per_partition = output_size_per_partition
        layer.orig_dtype = params_dtype
        device = extra_weight_attrs.get("device")
        weight = torch.nn.Parameter(torch.empty(output_size_per_partition, input_size_per_partition,
                                                dtype=torch.float8_e4m3fn, device=device), requires_grad=False)
        layer.register_parameter("weight", weight)
        assert self.quant_config.activation_scheme == "dynamic"
        scale = torch.empty((output_size_per_partition + block_n - 1) // block_n,
                            (input_size_per_partition + block_k - 1) // block_k, dtype=torch.float32, device=device)
        scale[:] = torch.finfo(torch.float32).min
        layer.register_parameter("weight_scale_inv", torch.nn.Parameter(scale, requires_grad=False))
        layer.register_parameter("input_scale", None)

    def process_weights_after_loading(self, layer: torch.nn.Module) -> None:
        # dense/layouts/fp8.py:213-232: block-quant weights are used as loaded (no requantisation, no transpose)
        layer.weight = torch.nn.Parameter(layer.weight.data.contiguous(), requires_grad=False)
        layer.weight_scale_inv = torch.nn.Parameter(layer.weight_scale_inv.data.contiguous(), requires_grad=False)

    def apply(self, layer: torch.nn.Module, x: torch.Tensor, bias: Optional[torch.Tensor] = None,
              block_scale: Optional[torch.Tensor] = None, output_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
        return self.w8a8_block_fp8_linear(input=x, weight=layer.weight,
                                          block_size=self.quant_config.weight_block_size,
                                          weight_scale=layer.weight_scale_inv, input_scale=block_scale, bias=bias,
                                          output_dtype=output_dtype)
