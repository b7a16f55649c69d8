#!/bin/bash
# Build attention A/B variants of the library (scripts/perf_attn.py / trace_attn.py pick them with SCAIL_LIB_VARIANT).
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $F $2 -o scail_b200/libscail_b200_$1.so scail_b200/csrc/api.cu -lcudart & }
rm -f scail_b200/libscail_b200_*.so
build p0  "-DSCAIL_ATT_POLY_MASK=0x0u"
build p12 "-DSCAIL_ATT_POLY_MASK=0x8080u"
build p25 "-DSCAIL_ATT_POLY_MASK=0x8888u"
build p37 "-DSCAIL_ATT_POLY_MASK=0x9249u"
build p50 "-DSCAIL_ATT_POLY_MASK=0xAAAAu"
build x25 "-DSCAIL_ATTN_EXPERIMENTS -DSCAIL_ATT_POLY_MASK=0x8888u"
wait
nvcc $F -o scail_b200/libscail_b200.so scail_b200/csrc/api.cu -lcudart
ls scail_b200/*.so
