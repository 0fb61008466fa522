#!/bin/bash
# ncu attribution of the final kernels (after the loop tunings): launch list + one full capture per dominant kernel.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r1h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1h_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_duo_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r1h_prof_sw_duo -f python bench.py --steps 1 --warmup 1 --workload sw --no-cpu-baseline > gpurun_out/r1h_ncu_sw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmm_duo_kernel --launch-skip 5 --launch-count 5 \
    -o gpurun_out/r1h_prof_hmm_duo -f python bench.py --steps 1 --warmup 1 --workload pairhmm --no-cpu-baseline > gpurun_out/r1h_ncu_hmm.log 2>&1
ls -la gpurun_out | grep r1h
