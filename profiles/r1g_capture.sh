#!/bin/bash
# Round-1 closing pass on one B200: full GPU test suite, default bench, reference arm, long-alignment bench.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r1g_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1g_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/r1g_bench.json 2> gpurun_out/r1g_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_bench_reference.json 2> gpurun_out/r1g_bench_reference.err
python bench.py --workload sw_long --steps 2 --warmup 1 > gpurun_out/r1g_bench_sw_long.json 2> gpurun_out/r1g_bench_sw_long.err
cat gpurun_out/r1g_pytest_gpu.log gpurun_out/r1g_smoke.log
