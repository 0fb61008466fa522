#!/bin/bash
# round 2, pass v: eight GPUs -- in-process multi-GPU tests, sw_long hand-off block sweep at 8 GPUs, the default bench under torchrun
mkdir -p gpurun_out
nvidia-smi -L | head -n 8
timeout 900 python -m pytest tests/test_multigpu_gpu.py -m gpu -q > gpurun_out/r2v_pytest_multigpu8.log 2>&1; echo "pytest exit $?"; tail -n 4 gpurun_out/r2v_pytest_multigpu8.log
: > gpurun_out/r2v_sw_long_b.jsonl
for cfg in "B=32" "B=16" "B=8" "B=16 R=2" "B=16 K=8"; do
  unset AGX_LONG_B AGX_LONG_R AGX_LONG_K
  for kv in $cfg; do case $kv in B=*) export AGX_LONG_B=${kv#B=};; R=*) export AGX_LONG_R=${kv#R=};; K=*) export AGX_LONG_K=${kv#K=};; esac; done
  timeout 300 python bench.py --workload sw_long --gpus 8 --steps 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        o=json.loads(l)['sw_long']; print(json.dumps({'cfg':'$cfg','ms':o['ms'],'ms_1gpu':o['ms_1gpu'],'speedup':o['speedup_vs_1gpu'],'kernel_ms_per_gpu':o['kernel_ms_per_gpu'],'score_ok':o['score_ok']}))" >> gpurun_out/r2v_sw_long_b.jsonl
done
unset AGX_LONG_B AGX_LONG_R AGX_LONG_K
cat gpurun_out/r2v_sw_long_b.jsonl
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 5 --warmup 3 \
    > gpurun_out/r2v_bench_n8.json 2> gpurun_out/r2v_bench_n8.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 4 gpurun_out/r2v_bench_n8.log; wc -c gpurun_out/r2v_bench_n8.json
