#!/bin/bash
mkdir -p gpurun_out
A="--trials 32 --steps 3 --warmup 3 --no-parity --no-long --no-bf16-leg --no-e2e"
for pdl in 0 1; do
  STGCN_PDL=$pdl timeout 400 python bench.py $A > gpurun_out/r2_c34_pdl$pdl.json 2> gpurun_out/r2_c34_pdl$pdl.err
  echo "== STGCN_PDL=$pdl" >> gpurun_out/r2_c34_rt.log
  STGCN_PDL=$pdl timeout 300 python tools/bench_rt.py --streams 17,64,256,1024,4096 --cuda-graph >> gpurun_out/r2_c34_rt.log 2>&1
done
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_c34_tests.log
echo done
