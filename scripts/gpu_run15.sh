set -x
cd $GRAFT_REPO_ROOT
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4
timeout 1200 python bench.py > gpurun_out/bench_r02_b.json 2> gpurun_out/bench_r02_b.err; tail -3 gpurun_out/bench_r02_b.err; grep '^{' gpurun_out/bench_r02_b.json | cut -c1-300
timeout 900 python bench.py --impl torchlib --steps 2 --warmup 2 > gpurun_out/bench_r02_torchlib.json 2> gpurun_out/bench_r02_torchlib.err; tail -3 gpurun_out/bench_r02_torchlib.err; grep '^{' gpurun_out/bench_r02_torchlib.json | cut -c1-300
timeout 600 python bench.py --latent 21x64x112 --steps 2 --warmup 2 --no-cpu-baseline --no-extras > gpurun_out/bench_r02_configB.json 2> gpurun_out/bench_r02_configB.err; grep '^{' gpurun_out/bench_r02_configB.json | cut -c1-300
