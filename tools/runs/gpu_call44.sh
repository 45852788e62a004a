#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c44_rt.log
: > $L
for pool in 1 0 1 0; do
  echo "== STGCN_RT_POOL=$pool" >> $L
  STGCN_RT_POOL=$pool timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --steps 400 >> $L 2>&1
done
timeout 600 python -m pytest tests -m gpu -x -q -k "rt or benchsize or top5" 2>&1 | tail -3 > gpurun_out/r2_c44_tests.log
echo done
