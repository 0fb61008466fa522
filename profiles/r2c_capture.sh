#!/bin/bash
# round 2, pass c: sweep of the software-pipelined long kernel; long-pair parity tests with it as the default
mkdir -p gpurun_out
python -m pytest tests/test_sw_gpu.py -m gpu -x -q -k "long or kbp or wave or stripe" > gpurun_out/r2c_pytest_long.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r2c_pytest_long.log
AGX_LIB_PATH=build/libagx_sweep.so timeout 400 python profiles/r2_long_sweep.py pipe_share > gpurun_out/r2c_long_sweep_pipe_share.jsonl 2>&1; echo "share exit $?"
AGX_LIB_PATH=build/libagx_sweep.so timeout 300 python profiles/r2_long_sweep.py pipe_full > gpurun_out/r2c_long_sweep_pipe_full.jsonl 2>&1; echo "full exit $?"
tail -n 4 gpurun_out/r2c_long_sweep_pipe_share.jsonl
