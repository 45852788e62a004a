#!/bin/bash
# round-2 GPU call 9: the new bench.py (parity, c1, rt cpu/e2e legs, CoST-GCN, long trial), 1 GPU
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_c9_bench.json 2> gpurun_out/r2_c9_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_c9_ref.json 2> gpurun_out/r2_c9_ref.err
STGCN_DEBUG=1 python bench.py --steps 1 > gpurun_out/r2_c9_refuse.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c9_refuse.log
echo done
