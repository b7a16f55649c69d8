#!/bin/bash
# 8-GPU box (gpurun --gpus 8): default layout (CFG-parallel x CP4) through torch.distributed and through the library's collective
cd ${GRAFT_REPO_ROOT:-.}
for native in 0 1; do
SCAIL_CP_NATIVE=$native timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2952$native bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu_native$native.json 2> gpurun_out/bench_8gpu_native$native.err; tail -2 gpurun_out/bench_8gpu_native$native.err | cut -c1-300; grep '^{' gpurun_out/bench_8gpu_native$native.json | cut -c1-200
done
