#!/bin/bash
# round 2, pass an: randomised end cell / start cell / CIGAR comparison with the oracle (class boundaries, ties, odd bytes, scorings)
mkdir -p gpurun_out
timeout 400 python profiles/align_fuzz.py 150 1 > gpurun_out/r2an_align_fuzz.jsonl 2> gpurun_out/r2an_align_fuzz.err; echo "fuzz exit $?"
tail -n 6 gpurun_out/r2an_align_fuzz.jsonl; tail -n 3 gpurun_out/r2an_align_fuzz.err
