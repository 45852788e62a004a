#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c35_rt.log
: > $L
timeout 300 python tools/bench_rt.py --streams 1,4,16 --cuda-graph >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1,4,16 --cuda-graph --math bf16 >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1,4,16 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> $L 2>&1
timeout 300 python tools/bench_rt.py --streams 1,16 --cuda-graph --graph imu_fogit_ABCD >> $L 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c35_tests.log
echo done
