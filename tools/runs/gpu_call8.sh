#!/bin/bash
# round-2 GPU call 8: two-stream pipeline of the graph-conv stage (GEMM of sub-chunk s+1 over LayerNorm of s)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c8_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e"
timeout 300 python bench.py $B32 > gpurun_out/r2_c8_overlap.json 2> gpurun_out/r2_c8_overlap.err
STGCN_OVERLAP=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c8_serial.json 2> gpurun_out/r2_c8_serial.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_c8_full.json 2> gpurun_out/r2_c8_full.err
echo done
