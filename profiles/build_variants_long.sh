#!/bin/bash
# builds build/libagx_<name>.so: libagx with sw_long.cu compiled under extra -D flags (kernel-tuning experiments)
set -e
cd "$(dirname "$0")/.."
C=accelerating-genomics_b200/csrc
mkdir -p build
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall --expt-relaxed-constexpr $flags -c $C/sw_long.cu -o build/sw_long_$name.o &
  names="$names $name"
done
wait
for name in $names; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libagx_$name.so $C/api.o $C/sw_kernels.o build/sw_long_$name.o $C/sw_parse.o $C/pairhmm_kernels.o $C/pairhmm_parse.o -lpthread
  echo built build/libagx_$name.so
done
