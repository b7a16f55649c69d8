set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -15
for v in r1 p0k3 p25k3 p37k3 p50k3 p25k2; do SDPA=$([ $v = r1 ] && echo 1 || echo 1) SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -2; done
N=48832 SCAIL_LIB_VARIANT=p25k3 timeout 300 python scripts/perf_attn.py 2>&1 | tail -2
N=48832 SCAIL_LIB_VARIANT=r1 SDPA=0 timeout 300 python scripts/perf_attn.py 2>&1 | tail -2
