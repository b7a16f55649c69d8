set -x
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "gemm" 2>&1 | tail -8
SCAIL_GEMM_CG=1 timeout 200 python scripts/perf_kernels.py 2>&1 | grep -E "^(qkv|out|fc1|fc2) "
SCAIL_GEMM_CG=2 timeout 200 python scripts/perf_kernels.py 2>&1 | grep -E "^(qkv|out|fc1|fc2) "
SCAIL_GEMM_CG=2 SCAIL_GEMM_GROUP_M=8 timeout 200 python scripts/perf_kernels.py 2>&1 | grep -E "^(qkv|out|fc1|fc2) "
SCAIL_GEMM_CG=2 SCAIL_GEMM_GROUP_M=24 timeout 200 python scripts/perf_kernels.py 2>&1 | grep -E "^(qkv|out|fc1|fc2) "
SCAIL_GEMM_CG=1 SCAIL_GEMM_GROUP_M=48 timeout 200 python scripts/perf_kernels.py 2>&1 | grep -E "^(qkv|out|fc1|fc2) "
