#!/bin/bash
# round 2, pass az: two-lane alignment shards (upload of chunk k+1 under the kernels of chunk k, results of chunk k under chunk k+1);
# file image with the aligned uniform segments
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py -q -m gpu -x > gpurun_out/r2az_pytest_align.log 2>&1; echo "align tests exit $?"; tail -n 3 gpurun_out/r2az_pytest_align.log
AGX_ALIGN_TRACE=1 timeout 900 python bench.py --no-sw-long --no-strong --no-gatk --sw-len "" --no-cpu-baseline > gpurun_out/r2az_bench.json 2> gpurun_out/r2az_bench.err; echo "bench exit $?"
grep "agx align\]" gpurun_out/r2az_bench.err | tail -n 24
python - <<'PY'
import json
for l in open('gpurun_out/r2az_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('sw value', round(d['value']), 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'], 'h2d only', d['e2e']['h2d_only_ms'], 'flat', d['e2e_flat']['ms_per_step'])
        a=d['sw_align']
        for k in ('ends','align'):
            print(k, round(a[k]['value']), round(a[k]['ms_per_step'],3), 'e2e', round(a[k]['e2e']['value']), round(a[k]['e2e']['ms_per_step'],2), a[k].get('dp_kernel_ms'), a[k].get('walk_kernel_ms'))
        print({k:v for k,v in a.items() if isinstance(v,bool)}, a.get('parity'))
PY
for c in 0 131072 524288; do AGX_ALIGN_CHUNK=$c timeout 300 python profiles/align_probe.py 1000000 150 ends+align 2>/dev/null | sed "s/^/chunk $c: /"; done
