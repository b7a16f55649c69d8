#!/bin/bash
# Run on the GPU box under gpurun (1 GPU).  Produces in gpurun_out/:
#   launches_<tag>.csv      every launch of one warm-up + one timed 14B step with its device time
#   attn_<tag>.ncu-rep      --set full capture of the self-attention kernel (b=2, 40 heads, N=27904)
#   gemm_<tag>.ncu-rep      --set full capture of the GEMM kernel (QKV projection shape)
#   rows_<tag>.ncu-rep      --set full capture of ln_modulate_kernel and rmsnorm_rope_kernel
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_fwd -s 1 -c 1 -f -o gpurun_out/attn_${TAG} \
    python scripts/perf_kernels.py > gpurun_out/ncu_attn_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 2 -c 1 -f -o gpurun_out/gemm_${TAG} \
    python scripts/perf_kernels.py > gpurun_out/ncu_gemm_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_modulate|rmsnorm_rope" -s 2 -c 2 -f -o gpurun_out/rows_${TAG} \
    python scripts/perf_kernels.py > gpurun_out/ncu_rows_${TAG}.log 2>&1
ls -la gpurun_out/
