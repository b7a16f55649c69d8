#!/bin/bash
# Build attention A/B variants of the library (scripts/perf_attn.py / trace_attn.py pick them with SCAIL_LIB_VARIANT).
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $F $2 -o scail_b200/libscail_b200_$1.so scail_b200/csrc/api.cu -lcudart & }
rm -f scail_b200/libscail_b200_*.so
build s4k3  "-DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_K_STAGES=3"
build s4k2  "-DSCAIL_ATT_P_SPLIT=4 -DSCAIL_ATT_K_STAGES=2"
build s3k3  "-DSCAIL_ATT_P_SPLIT=3 -DSCAIL_ATT_K_STAGES=3"
build s2k3  "-DSCAIL_ATT_P_SPLIT=2 -DSCAIL_ATT_K_STAGES=3"
build s3k3p19  "-DSCAIL_ATT_P_SPLIT=3 -DSCAIL_ATT_K_STAGES=3 -DSCAIL_ATT_POLY_MASK=0x8420u"
build x3k3  "-DSCAIL_ATTN_EXPERIMENTS -DSCAIL_ATT_P_SPLIT=3 -DSCAIL_ATT_K_STAGES=3"
wait
nvcc $F -o scail_b200/libscail_b200.so scail_b200/csrc/api.cu -lcudart
ls scail_b200/*.so
