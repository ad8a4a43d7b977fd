"""TEST / BENCH INFRASTRUCTURE ONLY.  Run pieces of the reference's own Python from /root/reference where that tree
exists (the build container; never on the GPU box): the function definitions are compiled from the reference's source
files as they stand, the way tests/golden/make_golden.py produces the golden vectors."""
from __future__ import annotations

import os

REFERENCE_PY = "/root/reference/python"


def _extract(path, names, class_name=None):
    """The reference's own function definitions, compiled from its source file as they stand (no import of the package:
    its module-level imports need the flashinfer fork and other things this image lacks)."""
    import ast
    from typing import Tuple

    import torch
    from torch.nn.functional import scaled_dot_product_attention
    tree = ast.parse(open(path).read())
    ns = {"torch": torch, "Tuple": Tuple, "scaled_dot_product_attention": scaled_dot_product_attention,
          "ceil_div": lambda a, b: (a + b - 1) // b}
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    for node in body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return [ns[n] for n in names]


def reference_cpu_path():
    """(decode attention loop, block-FP8 matmul, activation quantiser) of the reference itself --
    TorchNativeAttnBackend._run_sdpa_forward_decode (layers/attention/torch_native_backend.py:275-343),
    native_w8a8_block_fp8_matmul / native_per_token_group_quant_fp8 (python/sglang/test/test_block_fp8.py:15-40,89-141)
    -- or None where /root/reference is absent."""
    if not os.path.isdir(REFERENCE_PY):
        return None
    try:
        (method,) = _extract(os.path.join(REFERENCE_PY, "sglang/srt/layers/attention/torch_native_backend.py"),
                             ["_run_sdpa_forward_decode"], class_name="TorchNativeAttnBackend")
        quant, matmul = _extract(os.path.join(REFERENCE_PY, "sglang/test/test_block_fp8.py"),
                                 ["native_per_token_group_quant_fp8", "native_w8a8_block_fp8_matmul"])

        def decode(query, output, k_cache, v_cache, req_to_token, req_pool_indices, seq_lens, scaling=None,
                   enable_gqa=False, causal=False):
            return method(None, query, output, k_cache, v_cache, req_to_token, req_pool_indices, seq_lens,
                          scaling=scaling, enable_gqa=enable_gqa, causal=causal)
        return decode, matmul, quant
    except Exception:
        return None
