#!/bin/bash
# round-2 GPU call 2: fused graph-conv stage with split LN work items; staged RT state kernel + bf16 FIFO
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_c2_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline"
timeout 600 python bench.py $B32 > gpurun_out/r2_c2_b32.json 2> gpurun_out/r2_c2_b32.err
STGCN_RT_STREAM=0 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph > gpurun_out/r2_c2_rt_old.log 2>&1
timeout 300 python tools/bench_rt.py --streams 256,4096 --cuda-graph >> gpurun_out/r2_c2_rt_new.log 2>&1
timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph --graph imu_fogit_ABCD --math bf16 >> gpurun_out/r2_c2_rt_new.log 2>&1
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python bench.py $N1 > gpurun_out/r2_c2_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_gcnw|k_tcn|k_embed|k_pool' -c 60 --csv --log-file gpurun_out/r2_c2_launches32.csv python bench.py $N1 > gpurun_out/r2_c2_ncu.log 2>&1
timeout 300 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c2_plain_rt.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'k_gcnw|k_rt_|k_embed|k_pool|k_advance' -s 300 -c 40 --csv --log-file gpurun_out/r2_c2_launches_rt.csv python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c2_ncu_rt.log 2>&1
echo done
