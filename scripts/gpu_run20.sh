set -x
cd $GRAFT_REPO_ROOT
SCAIL_LIB_VARIANT=spec timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_scale_gpu.py -x -q -m gpu -k "attention" -s 2>&1 | grep -E "attention N=|passed|failed" | tail -5
for i in 1 2; do
SDPA=0 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
SDPA=0 SCAIL_LIB_VARIANT=spec timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
done
SCAIL_LIB_VARIANT=specx timeout 300 python scripts/trace_attn.py 2>&1 | tail -16
