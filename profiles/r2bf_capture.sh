#!/bin/bash
# round 2, pass bf: hmm_duo_kernel -- the odd row of an odd K in its own float2 table (bank conflicts of K = 5 / 7)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pairhmm_gpu.py tests/test_gatk_pin.py -q -m gpu -x > gpurun_out/r2bf_pytest_hmm.log 2>&1; echo "hmm tests exit $?"; tail -n 3 gpurun_out/r2bf_pytest_hmm.log
for i in 1 2; do
timeout 600 python bench.py --workload pairhmm --no-cpu-baseline --no-gatk > gpurun_out/r2bf_bench_$i.json 2> gpurun_out/r2bf_bench_$i.err; echo "bench exit $?"
python - <<PY
import json
for l in open('gpurun_out/r2bf_bench_$i.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d.get('pairhmm', d)
        print('pairhmm', round(p['value'],1), 'kernel_ms', p['roofline']['kernel_ms'], 'e2e', round(p['e2e']['value'],1))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:hmm_duo_kernel --launch-skip 5 --launch-count 5 --csv --log-file gpurun_out/r2bf_hmm_conflicts.csv python bench.py --steps 1 --warmup 1 --workload pairhmm --no-cpu-baseline --no-gatk > /dev/null 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(l for l in open('gpurun_out/r2bf_hmm_conflicts.csv') if l.startswith('"'))]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
