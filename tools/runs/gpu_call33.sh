#!/bin/bash
mkdir -p gpurun_out
A="--trials 32 --steps 3 --warmup 3 --no-parity --no-long --no-bf16-leg --no-e2e"
for pdl in 0 1; do
  STGCN_PDL=$pdl timeout 400 python bench.py $A > gpurun_out/r2_c33_pdl$pdl.json 2> gpurun_out/r2_c33_pdl$pdl.err
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c33_tests.log
echo done
