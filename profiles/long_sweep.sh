#!/bin/bash
# sweep stripe width K, publish granularity RB and chain formulation of sw_long_kernel
# (1 Mbp x 1 Mbp unless LEN is set).  usage: GPUS=2 KS="8 16" RBS="32" CHAINS="0 1" bash profiles/long_sweep.sh
LEN=${LEN:-1000000}; GPUS=${GPUS:-1}
for c in ${CHAINS:-0}; do for k in ${KS:-8 16 32}; do for rb in ${RBS:-32 64 128}; do
  AGX_LONG_CHAIN=$c AGX_LONG_K=$k AGX_LONG_RB=$rb python bench.py --workload sw_long --gpus $GPUS --long-len $LEN --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('K=$k RB=$rb chain=$c gpus=$GPUS len=$LEN  %.1f ms  %.0f GCUPS  score %d' % (d['ms_per_step'], d['value'], d['config']['score']))"
done; done; done
