set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_8gpu_cfg.json 2> gpurun_out/bench_r02_8gpu_cfg.err; tail -3 gpurun_out/bench_r02_8gpu_cfg.err; grep '^{' gpurun_out/bench_r02_8gpu_cfg.json | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 4 --warmup 3 --no-cpu-baseline --parallel cp > gpurun_out/bench_r02_8gpu_cp.json 2> gpurun_out/bench_r02_8gpu_cp.err; tail -3 gpurun_out/bench_r02_8gpu_cp.err; grep '^{' gpurun_out/bench_r02_8gpu_cp.json | cut -c1-400
