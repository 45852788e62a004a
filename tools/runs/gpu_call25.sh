#!/bin/bash
mkdir -p gpurun_out
N1="--trials 32 --steps 5 --warmup 3 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e --no-long --no-parity"
timeout 300 python bench.py $N1 > gpurun_out/r2_c25_n32.json 2> gpurun_out/r2_c25_n32.err
timeout 300 python bench.py $N1 --math bf16 > gpurun_out/r2_c25_n32_bf16.json 2> gpurun_out/r2_c25_n32_bf16.err
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c25_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-rt --no-long > gpurun_out/r2_c25_bench.json 2> gpurun_out/r2_c25_bench.err
echo done
