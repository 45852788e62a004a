#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "graph or windows or c1 or model_c1 or rt_" 2>&1 | tail -4 > gpurun_out/r2_c24_tests.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-rt --no-long --no-bf16-leg --no-e2e --no-parity > gpurun_out/r2_c24_c1.json 2> gpurun_out/r2_c24_c1.err
timeout 300 python tools/bench_rt.py --streams 1,256 --cuda-graph > gpurun_out/r2_c24_rt.log 2>&1
echo done
