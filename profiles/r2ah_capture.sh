#!/bin/bash
# round 2, pass ah: hmm_duo_kernel occupancy targets and branch-free steps per loop trip
mkdir -p gpurun_out
: > gpurun_out/r2ah_hmm_variants.jsonl
for v in f16 f16o2 f24 f32 f24o2 default; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  timeout 300 python bench.py --workload pairhmm --no-cpu-baseline --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(json.dumps({'variant':'$v','value':d['value'],'kernel_ms':r['kernel_ms'],'frac':r['frac']}))" >> gpurun_out/r2ah_hmm_variants.jsonl
done
cat gpurun_out/r2ah_hmm_variants.jsonl
