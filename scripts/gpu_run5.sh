set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -5
for v in s1p0 s4p0 s1p25 s2p25 s4p25 s4p37 s4p50; do SDPA=0 SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SDPA=1 SCAIL_LIB_VARIANT=s4p25 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
for m in 0 1; do SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=x4p25 timeout 300 python scripts/trace_attn.py 2>&1 | tail -16; done
for m in 0 1 5 6; do SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=x1p0 timeout 300 python scripts/trace_attn.py 2>&1 | tail -16; done
