set -x
cd $GRAFT_REPO_ROOT
for v in s4k3 s4k2 s3k3 s2k3 s3k3p19; do SDPA=0 SCAIL_LIB_VARIANT=$v timeout 300 python scripts/perf_attn.py 2>&1 | tail -1; done
SCAIL_LIB_VARIANT=x3k3 timeout 300 python scripts/trace_attn.py 2>&1 | tail -16
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/bench_r02_a.json 2> gpurun_out/bench_r02_a.err; tail -3 gpurun_out/bench_r02_a.err; cat gpurun_out/bench_r02_a.json
