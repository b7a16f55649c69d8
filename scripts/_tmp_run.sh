cd $GRAFT_REPO_ROOT
( time timeout 1200 python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 ) 2>&1 | tail -6 | cut -c1-900
