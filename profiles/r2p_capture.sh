#!/bin/bash
# round 2, pass p: alignment path (staged batches unrolled, packing from the state registers) -- tests, bench with the sw_align object
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_align_gpu.py tests/test_drivers_gpu.py -q -m gpu > gpurun_out/r2p_pytest_align.log 2>&1; echo "align tests exit $?"
tail -n 4 gpurun_out/r2p_pytest_align.log
timeout 300 python profiles/align_probe.py 800000 150 > gpurun_out/r2p_align_probe.jsonl 2> gpurun_out/r2p_align_probe.err; echo "probe exit $?"
cat gpurun_out/r2p_align_probe.jsonl
timeout 900 python bench.py --no-sw-long --no-strong --no-gatk --sw-len "" > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo "bench exit $?"
tail -n 5 gpurun_out/r2p_bench.err
python - <<'PY'
import json
for l in open('gpurun_out/r2p_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(json.dumps(d.get('sw_align'),indent=1)[:5000]); print(d.get('parity'))
PY
