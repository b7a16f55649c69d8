set -x
cd $GRAFT_REPO_ROOT
./scripts/umma_microbench 2>&1 | tee gpurun_out/umma_microbench.jsonl
for m in 0 1 5 6; do SDPA=0 SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=v2x timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
for m in 0 1 5 6; do SDPA=0 SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=r1x timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SCAIL_LIB_VARIANT=v2x timeout 300 python scripts/trace_attn.py 2>&1 | tail -30
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu -s 2>&1 | tail -25
