#!/bin/bash
# round 2, pass ap: row steps of sw_longr_kernel per loop trip (register moves of the rotating state)
mkdir -p gpurun_out
: > gpurun_out/r2ap_long_unroll.txt
for v in default lu2 lu4; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  for shape in "1000000 1000000" "125000 1000000"; do
    echo -n "$v: " >> gpurun_out/r2ap_long_unroll.txt
    REPS=2 timeout 120 python profiles/long_probe.py $shape 2>&1 | tail -n 1 >> gpurun_out/r2ap_long_unroll.txt
  done
done
cat gpurun_out/r2ap_long_unroll.txt
