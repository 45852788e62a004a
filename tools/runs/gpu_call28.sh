#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c28_rt.log
: > $L
for mode in 0 1 2; do for cap in 120000 240000; do
  echo "== mode $mode cap $cap" >> $L
  STGCN_RT_OVERLAP_MODE=$mode STGCN_RT_OVERLAP_SMEM=$cap timeout 300 python tools/bench_rt.py --streams 2048,4096 --cuda-graph >> $L 2>&1
done; done
echo "== off" >> $L
STGCN_RT_OVERLAP=0 timeout 300 python tools/bench_rt.py --streams 2048,4096 --cuda-graph >> $L 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "rt or benchsize or top5" 2>&1 | tail -4 > gpurun_out/r2_c28_tests.log
echo done
