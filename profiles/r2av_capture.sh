#!/bin/bash
# round 2, pass av: alignment entry points with the remembered device-memory figure -- tests, bench object inside the full default workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu > gpurun_out/r2av_pytest_align.log 2>&1; echo "align tests exit $?"; tail -n 3 gpurun_out/r2av_pytest_align.log
timeout 900 python bench.py --no-sw-long --no-strong --no-gatk --sw-len "" --no-cpu-baseline > gpurun_out/r2av_bench.json 2> gpurun_out/r2av_bench.err; echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/r2av_bench.json'):
    if l.startswith('{'):
        a=json.loads(l)['sw_align']
        for k in ('ends','align'):
            print(k, round(a[k]['value']), round(a[k]['ms_per_step'],3), 'e2e', round(a[k]['e2e']['value']), round(a[k]['e2e']['ms_per_step'],2), a[k].get('kernels_gcups'))
PY
