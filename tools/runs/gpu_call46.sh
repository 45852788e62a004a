#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "rt or benchsize or top5" 2>&1 | tail -3 > gpurun_out/r2_c46_tests.log
timeout 120 python tools/bench_rt.py --streams 4096 --cuda-graph --steps 300 > gpurun_out/r2_c46_rt.log 2>&1
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
STGCN_RT_OVERLAP=0 timeout 100 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r02_plain_rt.log 2>&1 &&
STGCN_RT_OVERLAP=0 timeout 300 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_rt_|k_embed|k_pool|k_advance' -s 230 -c 23 --csv --log-file gpurun_out/r02_launches_rt4096.csv python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r02_ncu_rt.log 2>&1
echo done
