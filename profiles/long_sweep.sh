#!/bin/bash
# sweep stripe width K, chain form and rows per step of the long-alignment kernels
# (1 Mbp x 1 Mbp unless COLS/ROWS are set).  usage: GPUS=2 KS="7 8 14" CHAINS="0 1" ROWSS="1 2" bash profiles/long_sweep.sh
COLS=${COLS:-1000000}; ROWS=${ROWS:-1000000}; GPUS=${GPUS:-1}
for r in ${ROWSS:-1 2}; do for c in ${CHAINS:-0 1}; do for k in ${KS:-7 8 14 27 32}; do
  AGX_LONG_ROWS=$r AGX_LONG_CHAIN=$c AGX_LONG_K=$k python profiles/long_probe.py $COLS $ROWS $GPUS | sed "s/^/rows_per_step=$r /"
done; done; done
