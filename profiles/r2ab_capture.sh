#!/bin/bash
# round 2, pass ab: start slack again after the deferred unpacking + ncu of the lone warp
mkdir -p gpurun_out
: > gpurun_out/r2ab_long.txt
for S in 0 1 2; do
  echo -n "slack=$S: " >> gpurun_out/r2ab_long.txt
  AGX_LONG_SLACK=$S REPS=2 timeout 120 python profiles/long_probe.py 125000 1000000 2>&1 | tail -n 1 >> gpurun_out/r2ab_long.txt
done
cat gpurun_out/r2ab_long.txt
REPS=1 AGX_LONG_K=7 AGX_LONG_R=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_longr_kernel -c 1 -f \
    -o gpurun_out/r2ab_longr_7_4 python profiles/long_probe.py 125000 200000 > gpurun_out/r2ab_ncu.log 2>&1; echo "ncu exit $?"
python profiles/summarize_ncu.py gpurun_out/r2ab_longr_7_4.ncu-rep > gpurun_out/r2ab_longr_7_4_ncu.txt; tail -n 12 gpurun_out/r2ab_longr_7_4_ncu.txt
