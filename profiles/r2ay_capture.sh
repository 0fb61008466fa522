#!/bin/bash
# round 2, pass ay: sw_score_file_image -- uniform upload segments + a short last one, scores copied home per region
mkdir -p gpurun_out
timeout 600 python profiles/seg_probe.py > gpurun_out/r2ay_seg_probe.jsonl 2> gpurun_out/r2ay_seg_trace.err; echo "probe exit $?"
cat gpurun_out/r2ay_seg_probe.jsonl; grep "agx" gpurun_out/r2ay_seg_trace.err | tail -n 16
timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_drivers_gpu.py tests/test_formats.py -q -m gpu -k "not long" > gpurun_out/r2ay_pytest.log 2>&1; echo "tests exit $?"; tail -n 3 gpurun_out/r2ay_pytest.log
