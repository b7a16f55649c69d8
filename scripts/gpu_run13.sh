set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_scale_gpu.py -x -q -m gpu -s 2>&1 | grep -E "attention N=|config-A block|VAE decode|3-step|passed|failed" | tee gpurun_out/r02_parity_scale.txt
bash scripts/ncu_profile.sh r02 2>&1 | tail -5
