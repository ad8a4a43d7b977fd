"""TEST INFRASTRUCTURE ONLY.  Compiles the PRODUCT's streaming kernels (sglang_fluentllm_b200/csrc/small_kernels.cu:
metadata, quantisers, fused prologue, de-quantiser, combine, activation quantiser) for the host with g++, against the
same execution-model shim the reference's kernels are built with (oracle/ref_build/shim/ref_shim.h), into
tests/hostsim/_build/libb200_small_kernels_host.so.  The CPU suite can then run the CUDA source of those kernels
against the reference's own kernel source (oracle/_ref) without a GPU.  It is a checker, not a code path: nothing in
the product package can load it, and the GPU tests remain the parity tests proper.

The kernel part of the file is compiled unmodified; the stream stops at the C-ABI section (host launchers), and the
PTX wrappers of csrc/ptx.cuh are replaced by the host stand-ins in tests/hostsim/shim/ptx.cuh.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libb200_small_kernels_host.so")
SRC = os.path.join(ROOT, "sglang_fluentllm_b200", "csrc", "small_kernels.cu")
REF_SHIM = os.path.join(ROOT, "oracle", "ref_build", "shim")


def build(force: bool = False) -> str:
    deps = [SRC, os.path.abspath(__file__), os.path.join(HERE, "driver.cpp"), os.path.join(HERE, "shim", "ptx.cuh"),
            os.path.join(HERE, "shim", "host_utils.h"), os.path.join(REF_SHIM, "ref_shim.h"),
            os.path.join(REF_SHIM, "ref_shim_impl.inc"), os.path.join(ROOT, "include", "b200_decode.h")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    text = open(SRC).read()
    cut = text.index("\nusing namespace b200;\n")          # the C-ABI entry points (host launchers) start here
    unit = "\n".join(["#undef __shared__", "#define __shared__",   # `extern __shared__ int32_t s_meta[]`
                      '#line 1 "%s"' % SRC, text[:cut], '#include "%s"' % os.path.join(HERE, "driver.cpp"), ""])
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-w", "-x", "c++", "-",
           "-include", os.path.join(REF_SHIM, "ref_shim.h"), "-I", os.path.join(HERE, "shim"),
           "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, "-o", LIB]
    r = subprocess.run(cmd, input=unit, text=True, capture_output=True, cwd=OUT_DIR)
    if r.returncode != 0:
        raise RuntimeError("host build of small_kernels.cu failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force=True))
