#!/bin/bash
# round-2 GPU call 17: two-stream pipeline of the graph-conv stage with shared-memory headroom (A/B)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_c17_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-long --no-parity"
timeout 300 python bench.py $B32 > gpurun_out/r2_c17_overlap.json 2> gpurun_out/r2_c17_overlap.err
STGCN_OVERLAP=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c17_serial.json 2> gpurun_out/r2_c17_serial.err
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-long > gpurun_out/r2_c17_full.json 2> gpurun_out/r2_c17_full.err
echo done
