set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_vae_gpu.py tests/test_scale_gpu.py -x -q -m gpu 2>&1 | tail -8
for v in w16p0 w16p12 w16p25 w16p37 w8p25; do SDPA=0 SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SDPA=1 SCAIL_LIB_VARIANT=w16p25 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
for m in 0 1; do SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=x16p25 timeout 300 python scripts/trace_attn.py 2>&1 | tail -16; done
