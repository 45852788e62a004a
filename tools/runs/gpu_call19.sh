#!/bin/bash
# round-2 GPU call 19: wider ring / LN-item sweep of the opt-in one-kernel graph-conv stage
mkdir -p gpurun_out
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-long --no-parity --no-bf16-leg"
for mb in 32 48 64; do for it in 1 2 4; do
  STGCN_GCNW_FUSE=1 STGCN_GCNW_RING_MB=$mb STGCN_GCNW_LN_ITERS=$it timeout 300 python bench.py $B32 > gpurun_out/r2_c19_sweep_${mb}_${it}.json 2> gpurun_out/r2_c19_sweep_${mb}_${it}.err
done; done
echo done
