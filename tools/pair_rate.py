"""cta_group::2 tcgen05.mma issue-rate probe (see umma_pair_rate_kernel).  Usage: python tools/pair_rate.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sglang_fluentllm_b200 import _lib
L = _lib.lab()
for mode, n, what in ((0, 128, "Q.K^T fp8 N=128"), (0, 256, "Q.K^T-form fp8 N=256"), (2, 128, "Q.K^T bf16 N=128"), (1, 256, "P.V N=256"),
                      (1, 128, "P.V-form N=128"), (3, 128, "MLA tile mix (16 qk8 + 4 qk16 + 8 pv)")):
    for iters in (64, 512):
        c = ctypes.c_float(0)
        assert L.b200_umma_pair_rate(mode, n, iters, ctypes.byref(c)) == 0, L.b200_lab_last_error()
        print("mode %d %-40s iters %4d: %7.1f cycles/MMA%s" % (mode, what, iters, c.value, "  (%.0f cycles/tile)" % (c.value * 28) if mode == 3 else ""))
