#!/bin/bash
# Round-1 final measurement pass on one B200 (run from the repo root under gpurun).
# Every number that is reported comes from a run WITHOUT a profiler; the ncu passes only attribute time.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r1f_pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r1f_bench.json 2> gpurun_out/r1f_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1f_bench_reference.json 2> gpurun_out/r1f_bench_reference.err
python bench.py --workload sw_long --steps 2 --warmup 1 > gpurun_out/r1f_bench_sw_long.json 2> gpurun_out/r1f_bench_sw_long.err
# launch list of the default bench command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r1f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1f_ncu_launch.log 2>&1
# one full capture per dominant kernel
ncu --set full --clock-control none --import-source on -k regex:sw_duo_kernel --launch-skip 4 --launch-count 1 \
    -o gpurun_out/r1f_prof_sw_duo -f python bench.py --steps 1 --warmup 1 --workload sw --no-cpu-baseline > gpurun_out/r1f_ncu_sw.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hmm_duo_kernel --launch-skip 5 --launch-count 5 \
    -o gpurun_out/r1f_prof_hmm_duo -f python bench.py --steps 1 --warmup 1 --workload pairhmm --no-cpu-baseline > gpurun_out/r1f_ncu_hmm.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:sw_long_kernel --launch-skip 1 --launch-count 1 \
    -o gpurun_out/r1f_prof_sw_long -f python profiles/long_probe.py 1000000 1000000 > gpurun_out/r1f_ncu_long.log 2>&1
ls -la gpurun_out
