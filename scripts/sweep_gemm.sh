#!/bin/bash
# Raster-group / L2-hint sweep of the CTA-pair GEMM at the four 14B shapes (40 back-to-back launches each, settled clocks).
cd ${GRAFT_REPO_ROOT:-.}
for sh in qkv out fc1 fc2; do
  for g in 6 8 12 16 24; do
    SHAPE=$sh SCAIL_GEMM_GROUP_M=$g SCAIL_GEMM_L2_HINTS=1 timeout 100 python scripts/perf_gemm.py | tail -1
  done
  SHAPE=$sh SCAIL_GEMM_GROUP_M=12 SCAIL_GEMM_L2_HINTS=0 timeout 100 python scripts/perf_gemm.py | tail -1
  SHAPE=$sh SCAIL_GEMM_CG=1 timeout 100 python scripts/perf_gemm.py | tail -1
  SHAPE=$sh CUBLAS=1 timeout 100 python scripts/perf_gemm.py | tail -1
done
