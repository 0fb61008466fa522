#!/bin/bash
# round 2, pass u: end cell of whole-GPU pairs; class table A/B of the inter-task kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu -k "whole_gpu or end_cells" > gpurun_out/r2u_pytest_long_ends.log 2>&1; echo "long ends tests exit $?"
tail -n 12 gpurun_out/r2u_pytest_long_ends.log
bash profiles/r2t_capture.sh
