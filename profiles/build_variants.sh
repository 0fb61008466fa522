#!/bin/bash
# builds build/libagx_<name>.so: libagx with sw_kernels.cu compiled under extra -D flags (kernel-tuning experiments;
# profiles/align_probe.py and friends pick one with AGX_LIB_PATH)
#   usage: profiles/build_variants.sh name "-DAGX_X=1 -DAGX_Y=2" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
C=accelerating-genomics_b200/csrc
mkdir -p build
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-O3,-Wall --expt-relaxed-constexpr $flags -c $C/sw_kernels.cu -o build/sw_kernels_$name.o &
  pids="$pids $!"; names="$names $name"
done
wait
for name in $names; do
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/libagx_$name.so $C/api.o build/sw_kernels_$name.o $C/sw_long.o $C/sw_parse.o $C/pairhmm_kernels.o $C/pairhmm_parse.o -lpthread
  echo built build/libagx_$name.so
done
