#!/bin/bash
# round 2, pass o: alignment kernels -- row-step unroll sweep (instruction cache), walk with diagonal batches
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu > gpurun_out/r2o_pytest_align.log 2>&1; echo "align tests exit $?"
tail -n 4 gpurun_out/r2o_pytest_align.log
: > gpurun_out/r2o_align_sweep.jsonl
for v in default u1 u2 u8; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  echo "{\"variant\": \"$v\"}" >> gpurun_out/r2o_align_sweep.jsonl
  timeout 300 python profiles/align_probe.py 800000 150 ends >> gpurun_out/r2o_align_sweep.jsonl 2>> gpurun_out/r2o_align_sweep.err
  timeout 300 python profiles/align_probe.py 800000 150 align >> gpurun_out/r2o_align_sweep.jsonl 2>> gpurun_out/r2o_align_sweep.err
done
cat gpurun_out/r2o_align_sweep.jsonl
