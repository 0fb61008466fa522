#!/bin/bash
# round 2, pass ao: four-lane sub-warps for pairs of up to 32 / 64 columns
mkdir -p gpurun_out
for v in default tab2; do
  if [ $v = default ]; then unset AGX_LIB_PATH; else export AGX_LIB_PATH=build/libagx_$v.so; fi
  timeout 600 python bench.py --no-sw-long --no-strong --no-gatk --hmm-batches 20 --no-cpu-baseline --no-align --sw-len "24,32,48,64" > gpurun_out/r2ao_bench_$v.json 2> gpurun_out/r2ao_bench_$v.err; echo "bench $v exit $?"
  python - <<PY
import json
for l in open('gpurun_out/r2ao_bench_$v.json'):
    if l.startswith('{'):
        d=json.loads(l)
        for r in d.get('sw_lengths',{}).get('lengths',[]): print('$v', r['len'], round(r['kernel_gcups']), round(r['alu_frac_executed'],3))
PY
done
if [ -f build/libagx_tab2.so ]; then AGX_LIB_PATH=build/libagx_tab2.so timeout 600 python -m pytest tests/test_sw_gpu.py tests/test_align_gpu.py -q -m gpu -k "not long" > gpurun_out/r2ao_pytest_tab2.log 2>&1; echo "tests tab2 exit $?"; tail -n 3 gpurun_out/r2ao_pytest_tab2.log; fi
