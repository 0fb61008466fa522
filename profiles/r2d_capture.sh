#!/bin/bash
# round 2, pass d: what one warp per SM sub-partition can issue (peaks "issue" mode) + ncu of the pipelined long kernel
mkdir -p gpurun_out
drivers/bin/agx_peaks 4 issue > gpurun_out/r2d_peaks_issue.jsonl 2>&1
export AGX_LIB_PATH=build/libagx_sweep.so REPS=1
AGX_LONG_K=7 AGX_LONG_R=4 AGX_LONG_DP4A=1 AGX_LONG_PIPE=1 AGX_LONG_B=32 ncu --set full --clock-control none --import-source on -k regex:sw_longp -c 1 -f \
    -o gpurun_out/r2d_longp_7_4 python profiles/long_probe.py 125000 200000 > gpurun_out/r2d_ncu_longp.log 2>&1
tail -n 2 gpurun_out/r2d_ncu_longp.log
