set -x
cd $GRAFT_REPO_ROOT
for v in s4k3 v3m0 v5m0 v5s100 v5s40 v4s100; do SDPA=0 SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SDPA=1 SCAIL_LIB_VARIANT=v5s100 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
for m in 0 1; do SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=v5s100x timeout 300 python scripts/trace_attn.py 2>&1 | tail -16; done
