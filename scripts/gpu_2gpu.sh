#!/bin/bash
# 2-GPU box (gpurun --gpus 2): context-parallel tests + the pure-CP bench through torch.distributed and through the library's collective
cd ${GRAFT_REPO_ROOT:-.}
timeout 600 python -m pytest tests/test_cp_gpu.py -x -q -m gpu -s 2>&1 | grep -E "^\[\(|passed|failed|Error|assert" | tail -8
for native in 0 1; do
SCAIL_CP_NATIVE=$native timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$native bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline --parallel cp > gpurun_out/bench_2gpu_cp_native$native.json 2> gpurun_out/bench_2gpu_cp_native$native.err; tail -2 gpurun_out/bench_2gpu_cp_native$native.err | cut -c1-300; grep '^{' gpurun_out/bench_2gpu_cp_native$native.json | cut -c1-200
done
