#!/bin/bash
# round 2, pass q: host timeline of sw_align_batch_flat (pinned buffers)
mkdir -p gpurun_out
AGX_ALIGN_TRACE=1 timeout 300 python profiles/align_probe.py 1000000 150 align > gpurun_out/r2q_align_probe.jsonl 2> gpurun_out/r2q_align_trace.err; echo "probe exit $?"
cat gpurun_out/r2q_align_probe.jsonl; tail -n 24 gpurun_out/r2q_align_trace.err
timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2q_align_probe2.jsonl 2>/dev/null; cat gpurun_out/r2q_align_probe2.jsonl
