#!/bin/bash
# round-2 GPU call 14: TMA-store epilogue of k_gcnw (A/B), full tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c14_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-long --no-parity"
timeout 300 python bench.py $B32 > gpurun_out/r2_c14_tma1.json 2> gpurun_out/r2_c14_tma1.err
STGCN_GCNW_TMA_OUT=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c14_tma0.json 2> gpurun_out/r2_c14_tma0.err
timeout 300 python tools/bench_rt.py --streams 256,4096 --cuda-graph > gpurun_out/r2_c14_rt.log 2>&1
STGCN_GCNW_TMA_OUT=0 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph >> gpurun_out/r2_c14_rt.log 2>&1
echo done
