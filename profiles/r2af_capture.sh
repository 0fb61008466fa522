#!/bin/bash
# round 2, pass af: closing pass with the final code -- two GPUs: multi-GPU tests; GPU 0: full suite, both bench arms, ncu of the alignment kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r2af_pytest_multigpu.log 2>&1; echo "multigpu tests exit $?"; tail -n 3 gpurun_out/r2af_pytest_multigpu.log
export CUDA_VISIBLE_DEVICES=0
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2af_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/r2af_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2af_smoke.log
( time python bench.py --impl reference > gpurun_out/r2af_bench_reference.json 2> gpurun_out/r2af_bench_reference.log ) 2>&1 | tail -n 3
( time python bench.py > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 3 gpurun_out/r2af_bench.log; wc -c gpurun_out/r2af_bench.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sw_duo_kernel|sw_walk_kernel" -c 2 -f \
    -o gpurun_out/r2af_prof_align python profiles/align_probe.py 250000 150 align > gpurun_out/r2af_ncu_align.log 2>&1; echo "ncu exit $?"
python profiles/summarize_ncu.py gpurun_out/r2af_prof_align.ncu-rep > gpurun_out/r2af_align_ncu.txt
REPS=1 ncu --set full --clock-control none --import-source on -k regex:sw_longr_kernel -c 1 -f \
    -o gpurun_out/r2af_prof_sw_long python profiles/long_probe.py 1000000 200000 > gpurun_out/r2af_ncu_long.log 2>&1; echo "ncu long exit $?"
python profiles/summarize_ncu.py gpurun_out/r2af_prof_sw_long.ncu-rep > gpurun_out/r2af_sw_long_ncu.txt
rm -f gpurun_out/r2af_prof_sw_long.ncu-rep
du -sh gpurun_out
