#!/bin/bash
# round 2, pass bp: the default bench under torchrun on eight GPUs with the final code
mkdir -p gpurun_out
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --warmup 3 \
    > gpurun_out/r2bp_bench_n8.json 2> gpurun_out/r2bp_bench_n8.log ) 2>&1 | tail -n 3
echo "bench exit $?"; grep -v "Warning\|pin = lambda\|^\*\*\*\|OMP_NUM" gpurun_out/r2bp_bench_n8.log | tail -n 5; wc -c gpurun_out/r2bp_bench_n8.json
python - <<'PY'
import json
for l in open('gpurun_out/r2bp_bench_n8.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), round(d['e2e']['value']), round(d['pairhmm']['value']), round(d['pairhmm']['e2e']['value']))
        print('strong sw', {k:(round(v['ms'],3), round(v['speedup_vs_1gpu'],2)) for k,v in d['strong']['sw'].items() if isinstance(v,dict)})
        print('strong hmm', {k:(round(v['ms'],3), round(v['speedup_vs_1gpu'],2)) for k,v in d['strong']['pairhmm'].items() if isinstance(v,dict)})
        s=d['sw_long']; print('sw_long', s['ms'], round(s['value']), s['speedup_vs_1gpu'], s['pipeline_efficiency'], s.get('score_ok'), s.get('score_matches_1gpu'))
PY
