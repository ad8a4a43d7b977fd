"""TEST INFRASTRUCTURE ONLY.  Recipe that compiles the reference's OWN kernel sources for the host, from where they lie
under /root/reference, into oracle/_ref/libflashmla_ref_cpu.so (git-ignored; it travels to the GPU box like the built
product library).  Nothing of the reference is copied into the repository: the sources are streamed through the
compiler.

  3rdparty/flashmla-fp8/csrc/kernels/get_mla_metadata.cu                 get_mla_metadata_kernel          (a-4)
  3rdparty/flashmla-fp8/csrc/kernels/ckv_fp8_per_token/cache_kernels.cu  quantize_single_kernel           (a-6)
                                                                          quant_and_save_cache_k_kernel    (a-7)
                                                                          dequantize_fused_indexed_kernel  (a-8)
  3rdparty/flashmla-fp8/csrc/kernels/ckv_fp8_per_token/mla_combine.cu    flash_fwd_mla_combine_kernel     (a-5)
  3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu                       per_token_group_quant_8bit_kernel (G-2)

The `__global__` functions are compiled unmodified by g++ against shim/ref_shim.h (threads, warp shuffles and barriers
on std::thread; NVIDIA's host implementations of bf16 / e4m3 from the CUDA toolkit headers).  The one thing the stream
filter removes is what no host compiler can parse: the `kernel<<<grid, block>>>(...)` launcher functions and their
explicit instantiations; oracle/ref_build/driver.cpp launches the kernels with the same grid / block shapes instead.
The attention kernel (splitkv_mla.cu: wgmma, TMA, CuTe) cannot be built this way.

Usage: python oracle/ref_build/build_ref.py [reference_root]   (default /root/reference)
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")
LIB = os.path.join(OUT_DIR, "libflashmla_ref_cpu.so")
KERNELS = "3rdparty/flashmla-fp8/csrc/kernels"


def strip_host_launchers(text: str) -> str:
    out, lines, i = [], text.split("\n"), 0
    while i < len(lines):
        ln = lines[i]
        if ln.startswith("#define INSTANTIATE_"):          # explicit instantiations of the launchers
            while not lines[i].startswith("#undef INSTANTIATE_"):
                i += 1
            i += 1
            continue
        if re.match(r"void (invoke_|run_|sgl_)\w+\(", ln):  # a launcher: drop it up to its closing brace
            if out and out[-1].startswith("template"):
                out.pop()
            while lines[i] != "}":
                i += 1
            i += 1
            continue
        if re.match(r"template void \w+<", ln):             # a launcher's explicit instantiation (one line)
            i += 1
            continue
        out.append(ln)
        i += 1
    res = "\n".join(out)
    assert "<<<" not in res, "a kernel launch survived the filter"
    return res


def _compile(unit: str, obj: str, include_dirs) -> None:
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-pthread", "-w", "-c", "-x", "c++", "-",
           "-include", os.path.join(HERE, "shim", "ref_shim.h")]
    for d in list(include_dirs) + [cuda_inc]:
        cmd += ["-I", d]
    cmd += ["-o", obj]
    r = subprocess.run(cmd, input=unit, text=True, capture_output=True, cwd=OUT_DIR)
    if r.returncode != 0:
        raise RuntimeError("reference host build failed (%s):\n%s" % (os.path.basename(obj), r.stderr[-4000:]))


def build(ref_root: str = "/root/reference", force: bool = False) -> str:
    kdir = os.path.join(ref_root, KERNELS)
    fdir = os.path.join(kdir, "ckv_fp8_per_token")
    srcs = [os.path.join(kdir, "get_mla_metadata.cu"), os.path.join(fdir, "cache_kernels.cu"),
            os.path.join(ref_root, "3rdparty/flashinfer/csrc/sgl_fp8_quantization.cu"), os.path.join(fdir, "mla_combine.cu")]
    mine = [os.path.join(HERE, "driver.cpp"), os.path.join(HERE, "driver_combine.cpp"), os.path.abspath(__file__)]
    for dirpath, _, files in os.walk(os.path.join(HERE, "shim")):
        mine += [os.path.join(dirpath, f) for f in files]
    for s in srcs:
        if not os.path.exists(s):
            raise FileNotFoundError(s)
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(f) for f in srcs + mine):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    shim = os.path.join(HERE, "shim")
    # translation unit 1: metadata + KV/q quantisers + activation quantiser (CUDA toolkit headers only; cutlass::ceil_div
    # comes from shim/cutlass_min so that the real CUTLASS tree is not needed here)
    unit1 = "\n".join([
        # get_mla_metadata.cu declares its dynamic shared memory `extern __shared__ int shared_mem[]`: a plain extern here
        "#undef __shared__", "#define __shared__",
        '#line 1 "%s"' % srcs[0], strip_host_launchers(open(srcs[0]).read()),
        # cache_kernels.cu: `__shared__ int cache_idx` is one variable per block
        "#undef __shared__", "#define __shared__ static",
        '#line 1 "%s"' % srcs[1], strip_host_launchers(open(srcs[1]).read()),
        # sgl_fp8_quantization.cu: its one shared array is declared `static __shared__`
        "#undef __shared__", "#define __shared__",
        '#line 1 "%s"' % srcs[2], strip_host_launchers(open(srcs[2]).read()),
        '#include "%s"' % os.path.join(HERE, "driver.cpp"), ""])
    obj1 = os.path.join(OUT_DIR, "ref_quant_meta.o")
    _compile(unit1, obj1, [shim, os.path.join(shim, "cutlass_min"), kdir, fdir])
    # translation unit 2: the split-KV combine kernel (its own params.h; CuTe tensors from the reference's vendored CUTLASS)
    unit2 = "\n".join([
        "#undef __shared__", "#define __shared__",     # `extern __shared__ float smem_buf[]`
        '#line 1 "%s"' % srcs[3], strip_host_launchers(open(srcs[3]).read()),
        '#include "%s"' % os.path.join(HERE, "driver_combine.cpp"), ""])
    obj2 = os.path.join(OUT_DIR, "ref_combine.o")
    _compile(unit2, obj2, [shim, fdir, os.path.join(ref_root, "3rdparty/flashmla-fp8/csrc/cutlass/include")])
    r = subprocess.run(["g++", "-shared", "-pthread", "-o", LIB, obj1, obj2], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference host link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(*(sys.argv[1:2]), force=True))
