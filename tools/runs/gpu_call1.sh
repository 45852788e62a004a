#!/bin/bash
# round-2 GPU call 1: fused graph-conv stage -- parity, A/B against the two-kernel form, full bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_c1_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-rt --no-cpu-baseline --no-bf16-leg"
timeout 300 python bench.py $B32 > gpurun_out/r2_c1_b32_fused.json 2> gpurun_out/r2_c1_b32_fused.err
STGCN_GCNW_FUSE=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c1_b32_unfused.json 2> gpurun_out/r2_c1_b32_unfused.err
timeout 300 python bench.py $B32 --math bf16 > gpurun_out/r2_c1_b32_fused_bf16.json 2> gpurun_out/r2_c1_b32_fused_bf16.err
STGCN_GCNW_FUSE=0 timeout 300 python bench.py $B32 --math bf16 > gpurun_out/r2_c1_b32_unfused_bf16.json 2> gpurun_out/r2_c1_b32_unfused_bf16.err
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c1_bench.json 2> gpurun_out/r2_c1_bench.err
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
timeout 300 python bench.py $N1 > gpurun_out/r2_c1_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_c1_launches32.csv python bench.py $N1 > gpurun_out/r2_c1_ncu.log 2>&1
echo done
