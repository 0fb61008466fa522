#!/bin/bash
# round 2, pass ai: register budget of the alignment modes of the inter-task kernel (CTAs of 128 threads per SM)
mkdir -p gpurun_out
: > gpurun_out/r2ai_align_mb.jsonl
for v in default mb3 mb5; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  echo "{\"variant\": \"$v\"}" >> gpurun_out/r2ai_align_mb.jsonl
  timeout 300 python profiles/align_probe.py 1000000 150 ends+align >> gpurun_out/r2ai_align_mb.jsonl 2>/dev/null
done
cat gpurun_out/r2ai_align_mb.jsonl
