#!/bin/bash
# round-2 GPU call 15: top-5 in the logits kernel; RT state kernels at 256 / 1024 / 4096 streams
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c15_tests.log
timeout 300 python tools/bench_rt.py --streams 64,256,1024,4096 --cuda-graph > gpurun_out/r2_c15_rt_stream.log 2>&1
STGCN_RT_STREAM=0 timeout 300 python tools/bench_rt.py --streams 64,256,1024,4096 --cuda-graph > gpurun_out/r2_c15_rt_update.log 2>&1
echo done
