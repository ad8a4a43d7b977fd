"""CPU oracle for the B200 quantised-decode hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package (``sglang_fluentllm_b200``).  Only ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it,
and there only as the checker / the reported CPU baseline.

Every function is a plain numpy / torch-CPU restatement of the reference's
algorithm and cites the reference file:line it follows (paths relative to
``/root/reference``).

Pinning status (see DESIGN.md "Oracle pinning"):
  * quantiser fall-back, indexed dequant, block-table allocator, torch-native
    SDPA decode, block-FP8 GEMM and 1x128 activation quantiser are pinned by
    golden vectors generated from the reference's own Python (tests/golden/).
  * rotary embedding (oracle/rope.py) is pinned bit-for-bit by the reference's
    DeepseekScalingRotaryEmbedding.forward_native (tests/golden/rope.npz).
  * the attention restatements (bf16, semantic FP8, page-wise faithful FP8) pass the
    reference's own kernel test -- its inputs, its fp32 ref_mla outputs and its
    relative_L1 < 5e-2 criterion (tests/golden/mla_test_reference.npz).
  * the metadata scheduler, the three KV/q quantiser kernels and the 1x128 activation
    quantiser kernel are pinned BIT FOR BIT by
    the reference's own CUDA sources compiled for the host and run on CPU threads
    (oracle/ref_build/ -> oracle/_ref/libflashmla_ref_cpu.so, oracle/ref_cpu.py;
    fixture tests/golden/ref_kernels_cpu.npz).
  * the per-token-FP8 MLA decode kernel exists in the reference only as sm_90a CUDA
    (wgmma / TMA / CuTe) that can neither execute on this box nor be compiled for the
    host: PARITY UNPINNED for its bit-level behaviour; its restatements are pinned at
    the reference's own test criterion (above).
"""
