#!/bin/bash
# round 2, pass bk: pairhmm_forward_batches_flat -- parts of a shard alternate between two lanes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pairhmm_gpu.py tests/test_gatk_pin.py tests/test_drivers_gpu.py -q -m gpu -x > gpurun_out/r2bk_pytest_hmm.log 2>&1; echo "hmm tests exit $?"; tail -n 3 gpurun_out/r2bk_pytest_hmm.log
timeout 600 python profiles/hmm_flat_probe.py > gpurun_out/r2bk_hmm_flat.jsonl 2> gpurun_out/r2bk_hmm_flat.err; echo "probe exit $?"; cat gpurun_out/r2bk_hmm_flat.jsonl; tail -n 3 gpurun_out/r2bk_hmm_flat.err
