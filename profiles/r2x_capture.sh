#!/bin/bash
# round 2, pass x: ncu of the storing kernel (MODE 2) and the traceback walk at their final form
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sw_duo_kernel|sw_walk_kernel" -c 2 -f \
    -o gpurun_out/r2x_prof_align2 python profiles/align_probe.py 250000 150 align > gpurun_out/r2x_ncu_align2.log 2>&1; echo "ncu exit $?"
python profiles/summarize_ncu.py gpurun_out/r2x_prof_align2.ncu-rep > gpurun_out/r2x_align2_ncu.txt; cat gpurun_out/r2x_align2_ncu.txt
