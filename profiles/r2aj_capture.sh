#!/bin/bash
# round 2, pass aj: rows per refill of the inter-task kernel (32 / 64 / 128)
mkdir -p gpurun_out
: > gpurun_out/r2aj_duo_ch.jsonl
for v in default ch64 ch128 default; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  timeout 300 python bench.py --no-sw-long --no-strong --no-gatk --hmm-batches 20 --no-cpu-baseline --no-align --sw-len "64,128,512" --steps 5 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print(json.dumps({'variant':'$v','value':d['value'],'kernel_ms':r['kernel_ms'],'frac':r['frac'],'lens':[(x['len'],round(x['kernel_gcups'])) for x in d['sw_lengths']['lengths']]}))" >> gpurun_out/r2aj_duo_ch.jsonl
done
cat gpurun_out/r2aj_duo_ch.jsonl
