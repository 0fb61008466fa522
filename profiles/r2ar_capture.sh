#!/bin/bash
# round 2, pass ar: sw_align bench object through the device-resident entry points (host clock, device synchronised)
mkdir -p gpurun_out
timeout 900 python bench.py --no-sw-long --no-strong --no-gatk --sw-len "" --no-cpu-baseline > gpurun_out/r2ar_bench.json 2> gpurun_out/r2ar_bench.err; echo "bench exit $?"; tail -n 3 gpurun_out/r2ar_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2ar_bench.json'):
    if l.startswith('{'):
        a=json.loads(l)['sw_align']
        for k in ('ends','align'):
            print(k, {x:a[k][x] for x in ('value','ms_per_step','dp_kernel_ms','gpu_launches') if x in a[k]}, a[k]['e2e']['value'], a[k].get('walk_kernel_ms'), a[k].get('kernels_gcups'))
        print({k:v for k,v in a.items() if isinstance(v,bool)})
PY
