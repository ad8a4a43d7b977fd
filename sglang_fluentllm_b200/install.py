"""The redirections of INTEGRATION.md as code: make the reference's own import statements land on the B200 modules.

``install()`` registers the mirrors under the names the reference imports its kernel packages by
(python/sglang/srt/layers/attention/flashmla_backend.py:14-15 ``flash_mla_fp8`` / ``flash_mla_swap``; that backend module
itself, imported by path in speculative/base_spec_worker.py:233-240;
python/sglang/srt/mem_cache/memory_pool.py:43-45; python/sglang/srt/layers/dense/gemms/fp8/deep_geem.py:6 ``deep_gemm``)
and must run before ``sglang.srt`` is imported.  ``patch_attention_initializer(module)`` puts the backend classes where
python/sglang/srt/model_executor/attn_initializer.py:25-29 would have imported the sm_90 ones from.  Nothing here
touches a GPU; a missing library still fails at the first kernel call (``_lib.lib()``)."""
from __future__ import annotations

import sys
from types import ModuleType
from typing import Dict

_ALIASES = ("flash_mla_fp8", "flash_mla_swap", "deep_gemm", "sglang.srt.layers.attention.flashmla_backend")
_saved: Dict[str, ModuleType] = {}


def install(force: bool = False) -> Dict[str, ModuleType]:
    """Register the mirrors in ``sys.modules``.  An already imported foreign module of the same name is replaced only
    with ``force=True`` (otherwise a RuntimeError says which one is in the way)."""
    from . import deep_gemm, flash_mla_fp8, flashmla_backend

    # the last entry serves `from sglang.srt.layers.attention.flashmla_backend import FlashMLAMultiStepDecodeBackend`
    # (speculative/base_spec_worker.py:233-240) and any other import of the reference's backend module by path
    targets = {"flash_mla_fp8": flash_mla_fp8, "flash_mla_swap": flash_mla_fp8, "deep_gemm": deep_gemm,
               "sglang.srt.layers.attention.flashmla_backend": flashmla_backend}
    for name, mod in targets.items():
        cur = sys.modules.get(name)
        if cur is not None and cur is not mod:
            if not force:
                raise RuntimeError("sglang_fluentllm_b200.install: module %r is already imported from %s" % (
                    name, getattr(cur, "__file__", "?")))
            _saved[name] = cur
        sys.modules[name] = mod
    return targets


def uninstall() -> None:
    """Undo ``install()`` (tests)."""
    from . import deep_gemm, flash_mla_fp8, flashmla_backend

    for name in _ALIASES:
        if sys.modules.get(name) in (flash_mla_fp8, deep_gemm, flashmla_backend):
            del sys.modules[name]
        if name in _saved:
            sys.modules[name] = _saved.pop(name)


def patch_attention_initializer(attn_initializer: ModuleType) -> None:
    """``attn_initializer`` = the imported ``sglang.srt.model_executor.attn_initializer``: define the two names its
    ``--attention-backend flashmla`` branch constructs (:394-395), which the reference only imports on sm_90 (:25-29)."""
    from . import flashmla_backend
    from .flashmla_backend import FlashMLABackend, FlashMLAMultiStepDecodeBackend

    attn_initializer.FlashMLABackend = FlashMLABackend
    attn_initializer.FlashMLAMultiStepDecodeBackend = FlashMLAMultiStepDecodeBackend
    # prefill (EXTEND / MIXED) stays on the reference's flashinfer MLA backend, the base class of its FlashMLABackend
    # (flashmla_backend.py:49,100-104): hand our class a factory for it when that module imports in this environment
    try:
        from sglang.srt.layers.attention.flashinfer_mla_backend import FlashInferMLAAttnBackend
    except Exception:  # fork-only flashinfer symbols missing: prefill then raises a clear error instead
        return
    flashmla_backend.set_prefill_backend_factory(
        lambda model_runner, skip_prefill, kv_indptr_buf, q_indptr_decode_buf: FlashInferMLAAttnBackend(
            model_runner, skip_prefill, kv_indptr_buf, q_indptr_decode_buf))
