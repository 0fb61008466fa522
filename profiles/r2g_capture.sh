#!/bin/bash
# round 2, pass g: second form of the software-pipelined long kernel (stage-wise slots, table prefetch)
mkdir -p gpurun_out
AGX_LIB_PATH=build/libagx_sweep.so timeout 400 python profiles/r2_long_sweep.py pipe2 > gpurun_out/r2g_long_sweep_pipe2.jsonl 2>&1; echo "sweep exit $?"
export AGX_LIB_PATH=build/libagx_sweep.so REPS=1
AGX_LONG_K=7 AGX_LONG_R=4 AGX_LONG_DP4A=1 AGX_LONG_PIPE=1 AGX_LONG_B=8 ncu --set full --clock-control none --import-source on -k regex:sw_longp -c 1 -f \
    -o gpurun_out/r2g_longp_7_4 python profiles/long_probe.py 125000 200000 > gpurun_out/r2g_ncu_longp.log 2>&1
tail -n 2 gpurun_out/r2g_ncu_longp.log
