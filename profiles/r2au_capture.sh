#!/bin/bash
# round 2, pass at: the device-resident alignment call inside bench.py (timeline)
mkdir -p gpurun_out
AGX_ALIGN_TRACE=1 timeout 900 python bench.py --no-sw-long --no-strong --no-gatk --sw-len "" --no-cpu-baseline --workload both > gpurun_out/r2au_bench.json 2> gpurun_out/r2au_bench.err; echo "bench exit $?"
grep "agx align" gpurun_out/r2au_bench.err | tail -n 16
python - <<'PY'
import json
for l in open('gpurun_out/r2au_bench.json'):
    if l.startswith('{'):
        a=json.loads(l)['sw_align']
        for k in ('ends','align'):
            print(k, a[k]['value'], a[k]['ms_per_step'])
PY
