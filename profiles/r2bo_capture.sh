#!/bin/bash
# round 2, pass bo: FINAL code on one GPU -- full GPU test suite, smoke, both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r2bo_pytest_gpu.log 2>&1; echo "gpu tests exit $?"; tail -n 3 gpurun_out/r2bo_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bo_smoke.log 2>&1; echo "smoke exit $?"; tail -n 1 gpurun_out/r2bo_smoke.log
( time python bench.py --impl reference > gpurun_out/r2bo_bench_reference.json 2> gpurun_out/r2bo_bench_reference.log ) 2>&1 | tail -n 3
( time python bench.py > gpurun_out/r2bo_bench.json 2> gpurun_out/r2bo_bench.log ) 2>&1 | tail -n 3
echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/r2bo_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('sw', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'h2d', round(d['e2e']['h2d_only_ms'],3), 'flat', round(d['e2e_flat']['value']), round(d['e2e_flat']['ms_per_step'],3), 'frac', round(d['roofline']['frac'],3))
        p=d['pairhmm']; print('hmm', round(p['value']), 'e2e', round(p['e2e']['value']), round(p['e2e']['ms_per_step'],3), 'flat', round(p['e2e_flat']['value']), round(p['e2e_flat']['ms_per_step'],3), 'kernel', p['roofline']['kernel_ms'], round(p['roofline']['frac'],3), round(p['roofline']['frac_algorithmic'],3))
        print('strong sw', {k:round(v['ms'],3) for k,v in d['strong']['sw'].items() if isinstance(v,dict)})
        print('strong hmm', {k:round(v['ms'],3) for k,v in d['strong']['pairhmm'].items() if isinstance(v,dict)})
        a=d['sw_align']
        for k in ('ends','align'): print(k, round(a[k]['value']), round(a[k]['ms_per_step'],3), 'e2e', round(a[k]['e2e']['value']), round(a[k]['e2e']['ms_per_step'],2))
        print({k:v for k,v in a.items() if isinstance(v,bool)}, a.get('parity'), d['parity']['sw_mismatches'], d['parity']['hmm_mismatches'], d['pairhmm_gatk']['pin']['ok'], d['sw_long']['score_ok'], d['sw_long']['ms'], d['cpu_baseline']['value'], d['pairhmm']['cpu_baseline']['value'])
for l in open('gpurun_out/r2bo_bench_reference.json'):
    if l.startswith('{'):
        d=json.loads(l); print('reference arm', d['value'], d['unit'], d.get('pairhmm',{}).get('value'))
PY
