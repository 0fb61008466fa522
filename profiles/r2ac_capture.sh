#!/bin/bash
# round 2, pass ac: when the next block's boundary entries are requested (eighths of a block) x start slack
mkdir -p gpurun_out
: > gpurun_out/r2ac_long.txt
for S in 0 1; do for Q in 0 2 4 6; do
  echo -n "slack=$S req=$Q/8: " >> gpurun_out/r2ac_long.txt
  AGX_LONG_SLACK=$S AGX_LONG_REQ=$Q REPS=2 timeout 120 python profiles/long_probe.py 125000 1000000 2>&1 | tail -n 1 >> gpurun_out/r2ac_long.txt
done; done
cat gpurun_out/r2ac_long.txt
