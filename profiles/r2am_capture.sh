#!/bin/bash
# round 2, pass am: alignment entry points after fusing the host passes (chunk cut + byte range, no offset rebasing)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py tests/test_drivers_gpu.py -q -m gpu > gpurun_out/r2am_pytest_align.log 2>&1; echo "align tests exit $?"; tail -n 4 gpurun_out/r2am_pytest_align.log
AGX_ALIGN_TRACE=1 timeout 300 python profiles/align_probe.py 1000000 150 ends+align > gpurun_out/r2am_align_probe.jsonl 2> gpurun_out/r2am_align_trace.err; echo "probe exit $?"
cat gpurun_out/r2am_align_probe.jsonl; tail -n 4 gpurun_out/r2am_align_trace.err
