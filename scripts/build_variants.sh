#!/bin/bash
# Build attention A/B variants of the library (scripts/perf_attn_variants.py picks them with SCAIL_LIB_VARIANT).
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC"
build() { nvcc $F $2 -o scail_b200/libscail_b200_$1.so scail_b200/csrc/api.cu -lcudart & }
build p0k3  "-DSCAIL_ATT_POLY_MASK=0x0u"
build p25k3 "-DSCAIL_ATT_POLY_MASK=0x8888u"
build p37k3 "-DSCAIL_ATT_POLY_MASK=0x9249u"
build p50k3 "-DSCAIL_ATT_POLY_MASK=0xAAAAu"
build p25k2 "-DSCAIL_ATT_POLY_MASK=0x8888u -DSCAIL_ATT_K_STAGES=2"
wait
ls -la scail_b200/*.so
