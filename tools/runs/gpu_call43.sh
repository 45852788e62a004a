#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c43_rt.log
: > $L
timeout 600 python -m pytest tests -m gpu -x -q -k "rt or benchsize or top5" 2>&1 | tail -4 > gpurun_out/r2_c43_tests.log
for pool in 0 1 0 1; do
  echo "== STGCN_RT_POOL=$pool" >> $L
  STGCN_RT_POOL=$pool timeout 300 python tools/bench_rt.py --streams 256,4096 --cuda-graph --steps 400 >> $L 2>&1
done
STGCN_RT_POOL=0 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --steps 400 --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
STGCN_RT_POOL=1 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --steps 400 --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
echo done
