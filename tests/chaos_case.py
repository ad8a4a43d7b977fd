"""Helper of test_gpu_parity.py::test_mla_pair_is_timing_independent (run as a subprocess: the trace flags are read once
per process).  Computes a few MLA decode cases `repeats` times and stores the outputs of every run."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import test_gpu_parity as T  # noqa: E402
from sglang_fluentllm_b200 import flash_mla_fp8 as fm  # noqa: E402

CASES = [(2, 128, 1, [9000, 4100], False), (40, 128, 1, [700 + 37 * i for i in range(40)], False),
         (3, 128, 2, [3000, 130, 1500], True), (5, 96, 1, [1, 64, 2000, 700, 5000], False),
         (80, 128, 1, [20000] + [100 + 3 * i for i in range(79)], False),
         (6, 16, 4, [3000, 130, 1500, 64, 900, 2500], True),  # M = 64: one CTA of each pair, half as many parts when tuned
         # peaky logits (spread ~6 bits): a third of the tiles move a row reference -- the drain / rescale / hand-off path
         (12, 128, 1, [4096 - 100 * i for i in range(12)], False, 0.1732)]


def main(out_path, repeats):
    res = {}
    for ci, case in enumerate(CASES):
        b, h, s_q, seqlens, causal = case[:5]
        scale = case[5] if len(case) > 5 else 192 ** -0.5
        c = T.make_case(fm, b, h, s_q, seqlens, nan_pad=False)
        sl = c["sl"].to(T.DEV)
        for sched in ("ref", "tuned"):
            meta, ns = fm.get_mla_metadata(sl, s_q * h, 1, fixed_overhead_blocks=None if sched == "ref" else fm.tuned_overhead_blocks(s_q * h),
                                           num_sm_parts=None if sched == "ref" else fm.tuned_num_sm_parts(s_q * h))
            a = (c["qn"], c["qr"], c["kl"], c["kr"], c["qs"], c["ks"], c["bt"].to(T.DEV), sl, 512, meta, ns)
            for r in range(repeats):
                o, l = fm.flash_mla_ckv_fp8_per_token(*a, causal=causal, softmax_scale=scale)
                torch.cuda.synchronize()
                res["o_%d_%s_%d" % (ci, sched, r)] = o.view(torch.int16).cpu().numpy()
                res["l_%d_%s_%d" % (ci, sched, r)] = l.cpu().numpy()
    np.savez(out_path, **res)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]))
