"""CPU: the C-ABI library builds, loads and exports every symbol include/b200_decode.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header="b200_decode.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built_lib():
    import __graft_entry__ as ge
    ge.build()
    return ge.LIB


def test_header_declares_the_boundary():
    syms = _declared_symbols()
    for s in ("b200_mla_get_metadata", "b200_mla_decode_ckv_fp8_per_token", "b200_mla_quantize_q",
              "b200_mla_quantize_and_cache_k", "b200_mla_dequantize_indexed", "b200_mla_combine",
              "b200_fp8_gemm_nt", "b200_act_quant_1x128", "b200_last_error"):
        assert s in syms


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    for s in _declared_symbols():
        assert hasattr(lib, s), "missing export: %s" % s


def test_python_binding_covers_every_symbol(built_lib):
    from sglang_fluentllm_b200 import _lib
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()
    _lib.lib()  # sets restype/argtypes for all of them; raises if one is missing


def test_probe_library_is_separate_from_the_product(built_lib):
    """The UMMA descriptor / issue-rate probes live in their own test-only library (include/b200_lab.h): the product
    library exports none of them and the probe library exports all of them."""
    import __graft_entry__ as ge
    from sglang_fluentllm_b200 import _lib
    lab_syms = _declared_symbols("b200_lab.h")
    assert "b200_umma_selftest" in lab_syms and sorted(_lib.LAB_EXPORTED_SYMBOLS) == lab_syms
    product, lab = ctypes.CDLL(built_lib), ctypes.CDLL(ge.LAB_LIB)
    for s in lab_syms:
        assert hasattr(lab, s), "missing export: %s" % s
        assert not hasattr(product, s), "probe symbol %s leaked into the product library" % s
    _lib.lab()


def test_pure_host_entry_points(built_lib):
    from sglang_fluentllm_b200 import _lib
    L = _lib.lib()
    # flash_api.cpp:41-43: sm_count / h_k / ceil(M / 64)
    assert L.b200_mla_num_sm_parts(16, 1, 148) == 148
    assert L.b200_mla_num_sm_parts(128, 1, 148) == 74
    assert L.b200_mla_num_sm_parts(64, 1, 132) == 132
    assert L.b200_mla_num_sm_parts(0, 1, 148) == 0
    # flash_api.cpp:365-368 sizing
    assert L.b200_mla_workspace_bytes(128, 1, 16, 148) == (128 + 148) * 16 * 513 * 4
    assert L.b200_last_error() is not None


def test_grouped_offset_layout_tile_bound(built_lib):
    """The offset-layout grouped GEMM launches b200_fp8_gemm_grouped_row_tiles(3, groups, m) row tiles; a group's tiles
    start at its first row rounded down to 4 (csrc/fp8_gemm.cu, GROUPED mode 3).  The grid must hold every tile the
    kernel's own enumeration produces, for any m_offset."""
    import random
    from sglang_fluentllm_b200 import _lib
    L = _lib.lib()
    rnd = random.Random(0)
    edge = [0, 0, 1, 2, 3, 4, 5, 125, 126, 127, 128, 129, 130, 253, 254, 255, 256]
    for _ in range(20000):
        groups = rnd.randint(1, 12)
        lens = [rnd.choice(edge + [rnd.randint(0, 600)]) for _ in range(groups)]
        off = [0]
        for n in lens:
            off.append(off[-1] + n)
        m = off[-1] + rnd.choice([0, 0, 1, 9, 200])
        need = sum((off[g + 1] - (off[g] & ~3) + 127) // 128 for g in range(groups) if off[g + 1] > off[g])
        assert need <= L.b200_fp8_gemm_grouped_row_tiles(3, groups, m), (lens, m)
    assert L.b200_fp8_gemm_grouped_row_tiles(1, 4, 1152) == 9
    assert L.b200_fp8_gemm_grouped_row_tiles(2, 4, 256) == 8


def test_sass_uses_blackwell_tensor_and_tma_paths(built_lib):
    """The product kernels must be tcgen05 / TMA code, not legacy mma.sync."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "UTCQMMA" in out and "UTCHMMA" in out, "no tcgen05.mma (UTC*MMA) in SASS"
    assert "UTMALDG" in out, "no TMA loads in SASS"
    assert "LDTM" in out, "no tcgen05.ld in SASS"
    assert "HMMA." not in out.replace("UTCHMMA", ""), "legacy mma.sync found"


def test_missing_library_fails_loudly(monkeypatch):
    from sglang_fluentllm_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libb200decode.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "sglang_fluentllm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), "%s imports oracle" % f
