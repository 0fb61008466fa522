#!/bin/bash
# round 2, pass ax: GATK-mode pin on the GPU; sw_score_file_image timeline and upload-segment knob
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gatk_pin.py -q -m gpu > gpurun_out/r2ax_pytest_gatk.log 2>&1; echo "gatk tests exit $?"; tail -n 3 gpurun_out/r2ax_pytest_gatk.log
timeout 600 python profiles/seg_probe.py > gpurun_out/r2ax_seg_probe.jsonl 2> gpurun_out/r2ax_seg_trace.err; echo "probe exit $?"
cat gpurun_out/r2ax_seg_probe.jsonl; grep "agx" gpurun_out/r2ax_seg_trace.err | tail -n 14
