#!/bin/bash
# round 2, pass z: start slack of the long kernel (blocks the left neighbour is allowed to get ahead) x block size
mkdir -p gpurun_out
: > gpurun_out/r2z_long_slack.txt
for shape in "125000 1000000" "1000000 1000000"; do
  for B in 8 16; do for S in 0 1 2 4; do
    echo -n "B=$B slack=$S: " >> gpurun_out/r2z_long_slack.txt
    AGX_LONG_B=$B AGX_LONG_SLACK=$S REPS=2 timeout 120 python profiles/long_probe.py $shape 2>&1 | tail -n 1 >> gpurun_out/r2z_long_slack.txt
  done; done
done
cat gpurun_out/r2z_long_slack.txt
