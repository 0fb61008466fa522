#!/bin/bash
# round 2, pass s: sw_long on the per-GPU share of an 8-GPU run with narrower stripes (two warps per scheduler)
mkdir -p gpurun_out
timeout 900 python profiles/r2_long_sweep.py smallk > gpurun_out/r2s_long_smallk.jsonl 2> gpurun_out/r2s_long_smallk.err; echo "sweep exit $?"
cat gpurun_out/r2s_long_smallk.jsonl; tail -n 3 gpurun_out/r2s_long_smallk.err
