#!/bin/bash
# VAE evidence: launch list of one decode + --set full capture of the row-tile conv kernel (96->96 @ 512^2 x 81 frames)
TAG=${1:-r01}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vae_${TAG}.csv \
    python scripts/perf_vae.py > gpurun_out/ncu_vae_${TAG}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv3d_row -s 60 -c 1 -f -o gpurun_out/conv_${TAG} \
    python scripts/perf_vae.py >> gpurun_out/ncu_vae_${TAG}.log 2>&1
