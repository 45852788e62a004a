#!/bin/bash
mkdir -p gpurun_out
A="--trials 32 --steps 2 --warmup 3 --no-parity --no-long --no-bf16-leg --no-e2e --no-cpu-baseline"
for r in 1 2; do
  STGCN_RT_OVERLAP=0 timeout 300 python bench.py $A > gpurun_out/r2_c29_off$r.json 2> gpurun_out/r2_c29_off$r.err
  timeout 300 python bench.py $A > gpurun_out/r2_c29_on$r.json 2> gpurun_out/r2_c29_on$r.err
done
echo done
