#!/bin/bash
# compute-sanitizer memcheck over the small-shape kernel tests (the reference has no sanitizer runs).
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 99 --log-file gpurun_out/memcheck.log \
    python -m pytest tests/test_kernels_gpu.py -q -x --timeout 600 \
    -k "gemm_bias and 128-256-64 or gemm_epilogues or ln_modulate and 256 or rmsnorm_rope and 256 or test_attention and 1-2-256-256 or test_attention and 300 or small_ops or patchify or gemm_cta_pair and 2048 or attention_partial and 4-200" \
    > gpurun_out/memcheck_pytest.log 2>&1
echo "sanitizer exit code $?" >> gpurun_out/memcheck_pytest.log
tail -3 gpurun_out/memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/memcheck.log | head -5
