#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c21_tests.log
for s in 1 0; do
  STGCN_WINDOWS_SHARE=$s timeout 300 python tools/bench_windows.py >> gpurun_out/r2_c21_windows.log 2>&1
  STGCN_WINDOWS_SHARE=$s timeout 300 python tools/bench_windows.py --W 300 --L 2000 >> gpurun_out/r2_c21_windows.log 2>&1
done
timeout 900 python bench.py --steps 5 --warmup 3 --no-rt --no-long > gpurun_out/r2_c21_bench.json 2> gpurun_out/r2_c21_bench.err
echo done
