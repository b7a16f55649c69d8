set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_vae_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in p0 p12 p25 p37 p50; do SDPA=0 SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SDPA=1 SCAIL_LIB_VARIANT=p25 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
N=48832 SDPA=1 SCAIL_LIB_VARIANT=p25 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
for m in 0 1; do SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=x25 timeout 300 python scripts/trace_attn.py 2>&1 | tail -16; done
timeout 300 python scripts/perf_kernels.py 2>&1 | tail -1
