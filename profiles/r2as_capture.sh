#!/bin/bash
# round 2, pass as: timeline of the device-resident alignment call
mkdir -p gpurun_out
AGX_ALIGN_TRACE=1 timeout 300 python profiles/align_dev_probe.py > gpurun_out/r2as_dev_probe.jsonl 2> gpurun_out/r2as_dev_trace.err; echo "probe exit $?"
cat gpurun_out/r2as_dev_probe.jsonl; tail -n 12 gpurun_out/r2as_dev_trace.err
