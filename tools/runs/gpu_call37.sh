#!/bin/bash
mkdir -p gpurun_out
STGCN_LIB=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_dbg.so STGCN_DEBUG=4 timeout 300 python tools/bench_rt.py --streams 1 --steps 30 > gpurun_out/r2_c37_dbg.log 2>&1
echo done
