#!/bin/bash
# round-2 GPU call 12: all-taps layer path (kernel 69, BatchNorm on tensor cores): tests + timings
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c12_tests.log
B="--trials 16 --steps 3 --warmup 2 --no-cpu-baseline --no-rt --no-e2e --no-long --no-bf16-leg"
timeout 300 python bench.py $B --norm BatchNorm > gpurun_out/r2_c12_bn_taps.json 2> gpurun_out/r2_c12_bn_taps.err
STGCN_TAPS=0 timeout 300 python bench.py $B --norm BatchNorm > gpurun_out/r2_c12_bn_simt.json 2> gpurun_out/r2_c12_bn_simt.err
echo done
