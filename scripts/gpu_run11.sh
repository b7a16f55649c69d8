set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m pytest tests/test_cp_gpu.py -x -q -m gpu -s 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_2gpu_cfg.json 2> gpurun_out/bench_r02_2gpu_cfg.err; tail -3 gpurun_out/bench_r02_2gpu_cfg.err; cat gpurun_out/bench_r02_2gpu_cfg.json | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --parallel cp > gpurun_out/bench_r02_2gpu_cp.json 2> gpurun_out/bench_r02_2gpu_cp.err; tail -3 gpurun_out/bench_r02_2gpu_cp.err; cat gpurun_out/bench_r02_2gpu_cp.json | cut -c1-1500
