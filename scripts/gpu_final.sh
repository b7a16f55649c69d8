set -x
cd $GRAFT_REPO_ROOT
bash scripts/gpu_validate.sh
bash scripts/sanitize.sh 2>&1 | tail -6
timeout 900 python scripts/sample_50step.py 2>&1 | tail -2
