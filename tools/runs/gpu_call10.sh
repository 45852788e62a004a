#!/bin/bash
# round-2 GPU call 10: LayerNorm tables in shared memory (temporal pair kernel) A/B; full GPU tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c10_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-e2e --no-long --no-parity"
timeout 300 python bench.py $B32 > gpurun_out/r2_c10_tab1.json 2> gpurun_out/r2_c10_tab1.err
STGCN_TCN_TAB_SMEM=0 timeout 300 python bench.py $B32 > gpurun_out/r2_c10_tab0.json 2> gpurun_out/r2_c10_tab0.err
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e --no-long --no-parity"
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__t_sector_hit_rate.pct"
timeout 300 python bench.py $N1 > gpurun_out/r2_c10_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_tcn|k_ln_stream|k_embed|k_pool' -s 34 -c 34 --csv --log-file gpurun_out/r2_c10_launches32.csv python bench.py $N1 > gpurun_out/r2_c10_ncu.log 2>&1
echo done
