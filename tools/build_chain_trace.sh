#!/bin/sh
# measurement build (tools/chain_trace.py): libdccn.so's objects + chain.cu compiled with the in-kernel timeline
cd "$(dirname "$0")/../dl_ofdm_b200" && nvcc -c -std=c++17 -O3 -lineinfo -DDCCN_CHAIN_TRACE -gencode arch=compute_100a,code=sm_100a \
  -Xcompiler -fPIC -o build/chain_trace.o csrc/chain.cu && \
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libdccn_chtrace.so build/dccn.o build/train.o build/chain_trace.o
