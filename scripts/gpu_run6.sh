set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_dit_gpu.py tests/test_vae_gpu.py -x -q -m gpu 2>&1 | tail -8
timeout 300 python scripts/perf_kernels.py 2>&1 | tail -1
