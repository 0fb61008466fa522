#!/bin/bash
# round 2, pass n: alignment path with the staged bulk stores (step-major matrices) -- tests, throughput, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu > gpurun_out/r2n_pytest_align.log 2>&1; echo "align tests exit $?"
tail -n 8 gpurun_out/r2n_pytest_align.log
timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2n_align_probe.jsonl 2> gpurun_out/r2n_align_probe.err; echo "probe exit $?"
cat gpurun_out/r2n_align_probe.jsonl; tail -n 5 gpurun_out/r2n_align_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sw_duo_kernel|sw_walk_kernel" -c 4 -f -o gpurun_out/r2n_align python profiles/align_probe.py 250000 150 align > gpurun_out/r2n_ncu.log 2>&1; echo "ncu exit $?"
tail -n 3 gpurun_out/r2n_ncu.log
