#!/bin/bash
# round-2 GPU call 6: fused graph-conv stage v4 (position-major LN items, larger ring); RT class times
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c6_tests.log
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt"
timeout 600 python bench.py $B32 > gpurun_out/r2_c6_b32.json 2> gpurun_out/r2_c6_b32.err
DBG=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_dbg.so
for bits in 2048 4096; do
  STGCN_LIB=$DBG STGCN_DEBUG=$bits timeout 300 python bench.py $B32 --no-bf16-leg --no-e2e > gpurun_out/r2_c6_b32_dbg$bits.json 2> gpurun_out/r2_c6_b32_dbg$bits.err
done
timeout 300 python tools/bench_rt.py --streams 4096 --profile > gpurun_out/r2_c6_rt_classes.log 2>&1
STGCN_RT_STREAM=0 timeout 300 python tools/bench_rt.py --streams 4096 --profile >> gpurun_out/r2_c6_rt_classes.log 2>&1
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python bench.py $N1 > gpurun_out/r2_c6_plain.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_tcn|k_embed|k_pool' -s 23 -c 23 --csv --log-file gpurun_out/r2_c6_launches32.csv python bench.py $N1 > gpurun_out/r2_c6_ncu.log 2>&1
echo done
