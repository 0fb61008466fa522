#!/bin/bash
# round 2, pass ak: the default bench under torchrun on two GPUs with the final code (strong object now carries the alignment entry point)
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 \
    > gpurun_out/r2ak_bench_n2.json 2> gpurun_out/r2ak_bench_n2.log ) 2>&1 | tail -n 3
echo "bench exit $?"; tail -n 4 gpurun_out/r2ak_bench_n2.log; wc -c gpurun_out/r2ak_bench_n2.json
python - <<'PY'
import json
for l in open('gpurun_out/r2ak_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['e2e']['value'], d['pairhmm']['value']); print(json.dumps(d['strong']['sw']['e2e_align'])); print(json.dumps(d['sw_long'])[:400])
PY
