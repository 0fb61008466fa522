#!/bin/bash
# round 2, pass f: two GPUs -- the in-process multi-GPU tests and the default bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -n 4
python -m pytest tests/test_multigpu_gpu.py -m gpu -x -q > gpurun_out/r2f_pytest_multigpu.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/r2f_pytest_multigpu.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r2f_bench_n2.json 2> gpurun_out/r2f_bench_n2.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 6 gpurun_out/r2f_bench_n2.log; wc -c gpurun_out/r2f_bench_n2.json
