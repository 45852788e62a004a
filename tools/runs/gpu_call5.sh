#!/bin/bash
# round-2 GPU call 5: ncu --set full of the fused graph-conv kernel and the RT state kernel; RT with 2-6 CTAs/SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "rt_" 2>&1 | tail -5 > gpurun_out/r2_c5_tests.log
timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph > gpurun_out/r2_c5_rt.log 2>&1
timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> gpurun_out/r2_c5_rt.log 2>&1
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
timeout 300 python bench.py $N1 > gpurun_out/r2_c5_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k k_gcnw -s 11 -c 1 -o gpurun_out/r2_c5_gcnw64 python bench.py $N1 > gpurun_out/r2_c5_ncu.log 2>&1
timeout 300 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c5_plain_rt.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k k_rt_stream -s 188 -c 1 -o gpurun_out/r2_c5_rtstream512 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c5_ncu_rt.log 2>&1
echo done
