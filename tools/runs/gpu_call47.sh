#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c47_rt.log
: > $L
for pool in 0 1 0 1; do
  echo "== STGCN_RT_POOL=$pool" >> $L
  STGCN_RT_POOL=$pool timeout 100 python tools/bench_rt.py --streams 4096 --cuda-graph --steps 300 >> $L 2>&1
done
echo done
