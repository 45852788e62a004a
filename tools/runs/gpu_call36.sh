#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c36_rt.log
: > $L
for st in 3 5 8; do
  echo "== weight ring stages $st" >> $L
  STGCN_LIB=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_s$st.so timeout 300 python tools/bench_rt.py --streams 1,8 --cuda-graph --steps 400 >> $L 2>&1
  STGCN_LIB=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_s$st.so timeout 300 python tools/bench_rt.py --streams 1 --cuda-graph --steps 400 --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
done
echo done
