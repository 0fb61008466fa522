#!/bin/bash
# round 2, pass b: ncu --set full of the long kernels on the per-GPU share shape (125 kbp x 200 kbp): the round-1
# two-rows kernel ("before") and the first R-rows kernel (K=7 R=4 lean dp4a, B=32)
mkdir -p gpurun_out
export AGX_LIB_PATH=build/libagx_sweep.so REPS=1
AGX_LONG_OLD=1 AGX_LONG_K=7 ncu --set full --clock-control none --import-source on -k regex:sw_long2 -c 1 -f -o gpurun_out/r2b_long2_old \
    python profiles/long_probe.py 125000 200000 > gpurun_out/r2b_ncu_old.log 2>&1
AGX_LONG_K=7 AGX_LONG_R=4 AGX_LONG_DP4A=1 AGX_LONG_CHAIN=0 AGX_LONG_B=32 ncu --set full --clock-control none --import-source on -k regex:sw_longr -c 1 -f \
    -o gpurun_out/r2b_longr_7_4 python profiles/long_probe.py 125000 200000 > gpurun_out/r2b_ncu_new.log 2>&1
tail -2 gpurun_out/r2b_ncu_old.log gpurun_out/r2b_ncu_new.log
ls -la gpurun_out/*.ncu-rep
