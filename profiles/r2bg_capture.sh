#!/bin/bash
# round 2, pass bg: bench.py's e2e against the probe's -- same call, timing variants
mkdir -p gpurun_out
timeout 600 python profiles/seg_probe2.py > gpurun_out/r2bg_seg_probe2.jsonl 2> gpurun_out/r2bg_seg_trace.err; echo "probe exit $?"
cat gpurun_out/r2bg_seg_probe2.jsonl; grep "agx\]" gpurun_out/r2bg_seg_trace.err | tail -n 14
