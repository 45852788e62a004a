#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c27_rt.log
echo "== overlap off" > $L
STGCN_RT_OVERLAP=0 timeout 300 python tools/bench_rt.py --streams 1024,2048,4096 --cuda-graph >> $L 2>&1
echo "== overlap on (cap 118K)" >> $L
timeout 300 python tools/bench_rt.py --streams 2048,4096 --cuda-graph >> $L 2>&1
echo "== overlap on (no cap)" >> $L
STGCN_RT_OVERLAP_SMEM=240000 timeout 300 python tools/bench_rt.py --streams 2048,4096 --cuda-graph >> $L 2>&1
echo "== overlap on from 1024 (cap 118K)" >> $L
STGCN_RT_OVERLAP=1024 timeout 300 python tools/bench_rt.py --streams 1024 --cuda-graph >> $L 2>&1
echo "== overlap on, eager" >> $L
timeout 300 python tools/bench_rt.py --streams 4096 >> $L 2>&1
echo "== imu bf16" >> $L
timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
STGCN_RT_OVERLAP=0 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "rt or benchsize or top5" 2>&1 | tail -4 > gpurun_out/r2_c27_tests.log
echo done
