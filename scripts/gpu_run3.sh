set -x
cd $GRAFT_REPO_ROOT
timeout 120 ./scripts/umma_microbench 2>&1 | tee gpurun_out/umma_microbench2.jsonl
timeout 900 python -m pytest tests/test_dit_gpu.py -x -q -m gpu 2>&1 | tail -15
