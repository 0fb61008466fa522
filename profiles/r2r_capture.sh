#!/bin/bash
# round 2, pass r: two GPUs -- multi-GPU tests (alignment shards, file image with one byte range per GPU), then the whole GPU suite on GPU 0
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r2r_pytest_multigpu.log 2>&1; echo "multigpu tests exit $?"
tail -n 6 gpurun_out/r2r_pytest_multigpu.log
CUDA_VISIBLE_DEVICES=0 timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2r_pytest_gpu.log 2>&1; echo "gpu tests exit $?"
tail -n 6 gpurun_out/r2r_pytest_gpu.log
CUDA_VISIBLE_DEVICES=0 timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2r_align_probe.jsonl 2>/dev/null; cat gpurun_out/r2r_align_probe.jsonl
