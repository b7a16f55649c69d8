set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_cp_gpu.py -x -q -m gpu -s -k "4gpu" 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_4gpu_cfg.json 2> gpurun_out/bench_r02_4gpu_cfg.err; tail -3 gpurun_out/bench_r02_4gpu_cfg.err; grep '^{' gpurun_out/bench_r02_4gpu_cfg.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --parallel cp > gpurun_out/bench_r02_4gpu_cp.json 2> gpurun_out/bench_r02_4gpu_cp.err; tail -3 gpurun_out/bench_r02_4gpu_cp.err; grep '^{' gpurun_out/bench_r02_4gpu_cp.json | cut -c1-400
