#!/bin/bash
# round-2 GPU call 18: warp-per-frame LayerNorm kernel (C = 64 default; C = 128 via STGCN_LN_WARP=800) A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_c18_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-long --no-parity"
timeout 300 python bench.py $B32 > gpurun_out/r2_c18_warp416.json 2> gpurun_out/r2_c18_warp416.err
STGCN_LN_WARP=800 timeout 300 python bench.py $B32 > gpurun_out/r2_c18_warp800.json 2> gpurun_out/r2_c18_warp800.err
STGCN_LN_WARP=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c18_warp0.json 2> gpurun_out/r2_c18_warp0.err
echo done
