#!/bin/bash
# round 2, pass t: class table of the inter-task kernel for 192 .. 512 columns (wider sub-warps, fewer columns per lane)
mkdir -p gpurun_out
for v in default tabB; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  timeout 600 python bench.py --no-sw-long --no-strong --no-gatk --hmm-batches 20 --no-cpu-baseline --no-align --sw-len "160,192,256,320,384,512,450-500,768" > gpurun_out/r2t_bench_$v.json 2> gpurun_out/r2t_bench_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
for l in open('gpurun_out/r2t_bench_$v.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for r in d.get('sw_lengths',{}).get('lengths',[]): print('$v', r['len'], round(r['kernel_gcups']), round(r['alu_frac_executed'],3))
PY
done
