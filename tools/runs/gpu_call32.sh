#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c32_rt.log
: > $L
for pdl in 0 1; do
  echo "== STGCN_PDL=$pdl" >> $L
  STGCN_PDL=$pdl timeout 300 python tools/bench_rt.py --streams 17,64,256,1024,4096 --cuda-graph >> $L 2>&1
  STGCN_PDL=$pdl timeout 300 python tools/bench_rt.py --streams 256,4096 >> $L 2>&1
  STGCN_PDL=$pdl timeout 300 python tools/bench_rt.py --streams 256,4096 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c32_tests.log
echo done
