#!/bin/bash
# round 2, pass bi: sw_score_batch_flat -- next chunk staged before the host blocks, eight chunks, results straight into a pinned array
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sw_gpu.py tests/test_drivers_gpu.py -q -m gpu -x -k "not long" > gpurun_out/r2bi_pytest_sw.log 2>&1; echo "sw tests exit $?"; tail -n 3 gpurun_out/r2bi_pytest_sw.log
for c in default 65536 250000 0; do
  if [ $c = default ]; then unset AGX_SW_CHUNK; else export AGX_SW_CHUNK=$c; fi
  timeout 300 python profiles/align_probe.py 1000000 150 score 2>/dev/null | sed "s/^/chunk $c: /"
done
unset AGX_SW_CHUNK
timeout 600 python bench.py --workload sw --no-cpu-baseline --no-align > gpurun_out/r2bi_bench.json 2> gpurun_out/r2bi_bench.err; echo "bench exit $?"
python - <<'PY'
import json
for l in open('gpurun_out/r2bi_bench.json'):
    if l.startswith('{'):
        d=json.loads(l)
        print('sw', round(d['value']), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'h2d', round(d['e2e']['h2d_only_ms'],3), 'flat', round(d['e2e_flat']['value']), round(d['e2e_flat']['ms_per_step'],3))
PY
