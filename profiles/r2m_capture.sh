#!/bin/bash
# round 2, pass m: alignment throughput after the chunk-bound fix + ncu of the MODE 2 kernel and the walk
mkdir -p gpurun_out
timeout 300 python profiles/align_probe.py 1000000 150 > gpurun_out/r2m_align_probe.jsonl 2> gpurun_out/r2m_align_probe.err; echo "probe exit $?"
cat gpurun_out/r2m_align_probe.jsonl; tail -n 5 gpurun_out/r2m_align_probe.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sw_duo_kernel|sw_walk_kernel" -c 6 -f -o gpurun_out/r2m_align python profiles/align_probe.py 250000 150 > gpurun_out/r2m_ncu.log 2>&1; echo "ncu exit $?"
tail -n 3 gpurun_out/r2m_ncu.log
