#!/bin/bash
mkdir -p gpurun_out
N1="--trials 32 --steps 5 --warmup 3 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e --no-long --no-parity"
timeout 300 python bench.py $N1 > gpurun_out/r2_c23_n32.json 2> gpurun_out/r2_c23_n32.err
timeout 600 python -m pytest tests -m gpu -x -q -k "model or layer or cost or tsplit" 2>&1 | tail -4 > gpurun_out/r2_c23_tests.log
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
N0="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e --no-long --no-parity"
timeout 600 ncu --metrics $M --clock-control none -k regex:'k_ln_stream|k_ln_warp' -s 12 -c 12 --csv --log-file gpurun_out/r2_c23_ln.csv python bench.py $N0 > gpurun_out/r2_c23_ncu.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --no-rt --no-long --no-bf16-leg --no-e2e --no-parity > gpurun_out/r2_c23_c1.json 2> gpurun_out/r2_c23_c1.err
echo done
