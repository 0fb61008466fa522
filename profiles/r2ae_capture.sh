#!/bin/bash
# round 2, pass ae: walks ordered by the size of the score
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py tests/test_drivers_gpu.py -q -m gpu > gpurun_out/r2ae_pytest_align.log 2>&1; echo "align tests exit $?"; tail -n 6 gpurun_out/r2ae_pytest_align.log
timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2ae_align_probe.jsonl 2>/dev/null; cat gpurun_out/r2ae_align_probe.jsonl
