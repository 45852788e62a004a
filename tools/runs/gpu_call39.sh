#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c39_rt.log
: > $L
timeout 300 python tools/bench_rt.py --streams 1,8,16,17,256,4096 --cuda-graph --steps 400 >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1,16,4096 --cuda-graph --steps 400 --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c39_tests.log
echo done
