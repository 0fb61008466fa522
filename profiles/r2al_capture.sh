#!/bin/bash
# round 2, pass al: host timeline of sw_align_batch_flat / sw_ends_batch_flat (pinned buffers) with the final code
mkdir -p gpurun_out
AGX_ALIGN_TRACE=1 timeout 300 python profiles/align_probe.py 1000000 150 ends+align > gpurun_out/r2al_align_probe.jsonl 2> gpurun_out/r2al_align_trace.err; echo "probe exit $?"
cat gpurun_out/r2al_align_probe.jsonl; tail -n 8 gpurun_out/r2al_align_trace.err
