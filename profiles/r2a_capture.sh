#!/bin/bash
# round 2, pass a: GPU tests with the R-rows long kernel as the product path, the variant sweep, dp4a pipe test
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -3 gpurun_out/r2a_pytest_gpu.log
drivers/bin/agx_peaks 4 dp4a > gpurun_out/r2a_peaks_dp4a.jsonl 2>&1; cat gpurun_out/r2a_peaks_dp4a.jsonl
AGX_LIB_PATH=build/libagx_sweep.so timeout 300 python profiles/r2_long_sweep.py share > gpurun_out/r2a_long_sweep_share.jsonl 2>&1; echo "share exit $?"
AGX_LIB_PATH=build/libagx_sweep.so timeout 300 python profiles/r2_long_sweep.py full > gpurun_out/r2a_long_sweep_full.jsonl 2>&1; echo "full exit $?"
sort -t'"' -k1,1 gpurun_out/r2a_long_sweep_share.jsonl | head -3
