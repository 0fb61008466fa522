#!/bin/bash
# round 2, pass aa: long kernel with entry fields taken apart at their use (early requests no longer wait on the spot)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sw_gpu.py -q -m gpu -k "long" > gpurun_out/r2aa_pytest_long.log 2>&1; echo "long tests exit $?"; tail -n 3 gpurun_out/r2aa_pytest_long.log
: > gpurun_out/r2aa_long.txt
for shape in "125000 1000000" "1000000 1000000" "125000 125000"; do
  for B in 8 16 32; do
    echo -n "B=$B: " >> gpurun_out/r2aa_long.txt
    AGX_LONG_B=$B REPS=2 timeout 120 python profiles/long_probe.py $shape 2>&1 | tail -n 1 >> gpurun_out/r2aa_long.txt
  done
done
cat gpurun_out/r2aa_long.txt
