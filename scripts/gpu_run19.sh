set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "attention" 2>&1 | tail -3
timeout 600 python -m pytest tests/test_cp_gpu.py -x -q -m gpu -s 2>&1 | grep -E "^\[\(|passed|failed|Error" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_4gpu_cfg_lf.json 2> gpurun_out/bench_r02_4gpu_cfg_lf.err; tail -3 gpurun_out/bench_r02_4gpu_cfg_lf.err; grep '^{' gpurun_out/bench_r02_4gpu_cfg_lf.json | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --parallel cp > gpurun_out/bench_r02_4gpu_cp_lf.json 2> gpurun_out/bench_r02_4gpu_cp_lf.err; tail -3 gpurun_out/bench_r02_4gpu_cp_lf.err; grep '^{' gpurun_out/bench_r02_4gpu_cp_lf.json | cut -c1-300
