#!/bin/bash
# round 2, pass bn: two GPUs -- alignment shards hand their runs over in pinned memory, offsets written in place; multi-GPU tests, default bench under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q -m gpu > gpurun_out/r2bn_pytest_multigpu.log 2>&1; echo "multigpu tests exit $?"; tail -n 3 gpurun_out/r2bn_pytest_multigpu.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r2bn_bench_n2.json 2> gpurun_out/r2bn_bench_n2.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 4 gpurun_out/r2bn_bench_n2.log; wc -c gpurun_out/r2bn_bench_n2.json
python - <<'PY'
import json
for l in open('gpurun_out/r2bn_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['pairhmm']['value'], d['pairhmm']['e2e']['value'])
        print(json.dumps(d['strong']['sw']['e2e_align'])); print(json.dumps(d['strong']['sw']['e2e'])); print(json.dumps(d['sw_long'])[:400])
PY
