set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -4
for v in product v3b v5b v5bp0 v5bp37; do if [ $v = product ]; then unset SCAIL_LIB_VARIANT; else export SCAIL_LIB_VARIANT=$v; fi; SDPA=0 timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SDPA=1 SCAIL_LIB_VARIANT=v5b timeout 300 python scripts/perf_attn.py 2>&1 | tail -1
SCAIL_LIB_VARIANT=v5bx timeout 300 python scripts/trace_attn.py 2>&1 | tail -16
SCAIL_LIB_VARIANT=v3bx timeout 300 python scripts/trace_attn.py 2>&1 | tail -16
