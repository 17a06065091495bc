#!/bin/sh
# debug build of the library with the in-kernel timeline enabled (not used by the product or the tests)
cd "$(dirname "$0")/.." && nvcc -shared -Xcompiler -fPIC -std=c++17 -O3 -lineinfo -DDCCN_TRACE \
  -gencode arch=compute_100a,code=sm_100a -o dl_ofdm_b200/libdccn_trace.so dl_ofdm_b200/csrc/dccn.cu dl_ofdm_b200/csrc/train.cu dl_ofdm_b200/csrc/chain.cu
