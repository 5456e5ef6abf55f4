#!/bin/sh
# builds tools/tc_timing (phase-timestamp harness for the tcgen05 GEMM)
cd "$(dirname "$0")/.." || exit 1
C=numpy_transformer_b200/csrc
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -DNT_TC_TIMING \
  tools/tc_timing.cu $C/runtime.cu $C/gemm_simt.cu -o tools/tc_timing
