#!/bin/bash
# round-2 GPU call 7: L2 eviction hints + ring / LN-item sweep for the fused graph-conv stage; CoST-GCN tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c7_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-bf16-leg"
for mb in 12 24 40; do for it in 4 7 13; do
  STGCN_GCNW_RING_MB=$mb STGCN_GCNW_LN_ITERS=$it timeout 300 python bench.py $B32 > gpurun_out/r2_c7_sweep_${mb}_${it}.json 2> gpurun_out/r2_c7_sweep_${mb}_${it}.err
done; done
STGCN_GCNW_FUSE=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c7_unfused.json 2> gpurun_out/r2_c7_unfused.err
timeout 300 python bench.py $B32 --math bf16 > gpurun_out/r2_c7_fused_bf16.json 2> gpurun_out/r2_c7_fused_bf16.err
STGCN_GCNW_FUSE=0 timeout 300 python bench.py $B32 --math bf16 > gpurun_out/r2_c7_unfused_bf16.json 2> gpurun_out/r2_c7_unfused_bf16.err
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python bench.py $N1 > gpurun_out/r2_c7_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_tcn|k_embed|k_pool' -s 23 -c 23 --csv --log-file gpurun_out/r2_c7_launches32.csv python bench.py $N1 > gpurun_out/r2_c7_ncu.log 2>&1
echo done
