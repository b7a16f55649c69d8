set -x
cd $GRAFT_REPO_ROOT
M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__cycles_elapsed.avg.per_second,gpu__time_duration.sum,lts__t_sector_hit_rate.pct"
run() { echo "== $*"; env "$@" ITERS=2 timeout 120 ncu $M --clock-control none -k regex:gemm_bf16 -s 3 -c 1 python scripts/perf_gemm.py 2>&1 | grep -E "dram__|per_second|duration|hit_rate" | awk '{print $1, $(NF-1), $NF}'; env "$@" timeout 120 python scripts/perf_gemm.py | tail -1; }
for sh in qkv out; do
run SHAPE=$sh SCAIL_GEMM_CG=1
run SHAPE=$sh SCAIL_GEMM_CG=2 SCAIL_GEMM_L2_HINTS=0
run SHAPE=$sh SCAIL_GEMM_CG=2 SCAIL_GEMM_L2_HINTS=1
run SHAPE=$sh SCAIL_GEMM_CG=2 SCAIL_GEMM_L2_HINTS=2
run SHAPE=$sh SCAIL_GEMM_CG=2 SCAIL_GEMM_L2_HINTS=1 SCAIL_GEMM_GROUP_M=16
run SHAPE=$sh SCAIL_GEMM_CG=2 SCAIL_GEMM_L2_HINTS=1 SCAIL_GEMM_GROUP_M=8
done
SHAPE=qkv CUBLAS=1 timeout 120 python scripts/perf_gemm.py | tail -1
SHAPE=out CUBLAS=1 timeout 120 python scripts/perf_gemm.py | tail -1
SHAPE=fc2 CUBLAS=1 timeout 120 python scripts/perf_gemm.py | tail -1
