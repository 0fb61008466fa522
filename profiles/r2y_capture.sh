#!/bin/bash
# round 2, pass y: where a lone warp of sw_longr_kernel<7,4> loses its cycles (per-GPU share of an 8-GPU run)
mkdir -p gpurun_out
REPS=1 AGX_LONG_K=7 AGX_LONG_R=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sw_longr_kernel -c 1 -f \
    -o gpurun_out/r2y_longr_7_4 python profiles/long_probe.py 125000 200000 > gpurun_out/r2y_ncu.log 2>&1; echo "ncu exit $?"
tail -n 2 gpurun_out/r2y_ncu.log
python profiles/summarize_ncu.py gpurun_out/r2y_longr_7_4.ncu-rep > gpurun_out/r2y_longr_7_4_ncu.txt; grep -v "l1tex\|fp64" gpurun_out/r2y_longr_7_4_ncu.txt
