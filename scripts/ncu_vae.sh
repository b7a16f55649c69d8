ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vae.csv python scripts/perf_vae.py > gpurun_out/ncu_vae.log 2>&1
