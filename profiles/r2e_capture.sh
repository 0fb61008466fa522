#!/bin/bash
# round 2, pass e: the default bench line (all sections) on one GPU
mkdir -p gpurun_out
nproc
( time python bench.py > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 5 gpurun_out/r2e_bench.log; wc -c gpurun_out/r2e_bench.json
