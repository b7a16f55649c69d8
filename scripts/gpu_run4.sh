set -x
cd $GRAFT_REPO_ROOT
for v in r1m0 r1m1 r1m2; do
  for m in 0 1; do SDPA=0 SCAIL_ATTN_DEBUG=$m SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
done
for v in r1m0 r1m1 r1m2; do SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_kernels.py 2>&1 | tail -1; done
