#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c41_rt.log
: > $L
timeout 600 python -m pytest tests -m gpu -x -q -k "rt" 2>&1 | tail -4 > gpurun_out/r2_c41_tests.log
timeout 300 python tools/bench_rt.py --streams 1,8,14 --cuda-graph --steps 400 >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1,14 --cuda-graph --steps 400 --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1 --cuda-graph --steps 400 --math bf16 >> $L 2>&1
STGCN_LIB=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_dbg.so STGCN_DEBUG=4 timeout 300 python tools/bench_rt.py --streams 1 --steps 30 2>&1 | grep rt_small | tail -1 >> $L
echo done
