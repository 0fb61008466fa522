#!/bin/bash
# round 2, pass l: alignment path after the class-offset fix
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu > gpurun_out/r2l_pytest_align.log 2>&1; echo "align tests exit $?"
tail -n 15 gpurun_out/r2l_pytest_align.log
timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2l_align_probe.jsonl 2> gpurun_out/r2l_align_probe.err; echo "probe exit $?"
cat gpurun_out/r2l_align_probe.jsonl; tail -n 5 gpurun_out/r2l_align_probe.err
