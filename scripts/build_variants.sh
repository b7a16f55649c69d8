#!/bin/bash
# Build attention A/B variants of the library for scripts/perf_attn.py / trace_attn.py (SCAIL_LIB_VARIANT=<tag> picks one).
# Knobs: -DSCAIL_ATT_P_SPLIT={1,2,3,4}  -DSCAIL_ATT_POLY_MASK=0x....u  -DSCAIL_ATT_K_STAGES={2,3}  -DSCAIL_MBAR_MODE={0,1,2}
#        -DSCAIL_ATTN_EXPERIMENTS (clock64 traces + SCAIL_ATTN_DEBUG ablation modes)
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $F $2 -o scail_b200/libscail_b200_$1.so scail_b200/csrc/api.cu -lcudart & }
rm -f scail_b200/libscail_b200_*.so
build s1p0  "-DSCAIL_ATT_P_SPLIT=1 -DSCAIL_ATT_POLY_MASK=0x0u"
build s4p0  "-DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_POLY_MASK=0x0u"
build s4p25 "-DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_POLY_MASK=0x8888u"
build s4p37 "-DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_POLY_MASK=0x9249u"
build x4p25 "-DSCAIL_ATTN_EXPERIMENTS -DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_POLY_MASK=0x8888u"
wait
ls scail_b200/*.so
