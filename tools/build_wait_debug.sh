#!/bin/bash
# Debug build of the product library with -DB200_WAIT_DEBUG (see csrc/ptx.cuh: wait_gave_up, B200_BEACON) for tools/chaos_debug.py.
# Overwrites sglang_fluentllm_b200/_C/libb200decode.so: run `python __graft_entry__.py` afterwards to restore the product build.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd $ROOT/sglang_fluentllm_b200/csrc
OUT=$ROOT/sglang_fluentllm_b200/_C
for f in host_utils small_kernels fp8_gemm mla_decode mla_decode_pair mla_decode_bf16 p2p_comm; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -DB200_WAIT_DEBUG -I$ROOT/include -I. -c $f.cu -o /tmp/dbg_$f.o &
done
wait
nvcc -shared -o $OUT/libb200decode.so /tmp/dbg_host_utils.o /tmp/dbg_small_kernels.o /tmp/dbg_fp8_gemm.o /tmp/dbg_mla_decode.o /tmp/dbg_mla_decode_pair.o /tmp/dbg_mla_decode_bf16.o /tmp/dbg_p2p_comm.o -lcuda -lcudart
