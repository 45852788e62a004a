#!/bin/bash
# round-2 measurement call: tests, full 1-GPU bench, launch lists (ST-GCN chunk, RT step), ncu --set full summary
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_tests.log
timeout 300 python tools/bench_rt.py --streams 256 --profile > gpurun_out/r02_rt256_classes.log 2>&1
timeout 300 python tools/bench_rt.py --streams 64,256,1024,4096 --cuda-graph > gpurun_out/r02_rt_sizes.log 2>&1
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e --no-long --no-parity"
M="gpu__time_duration.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum"
timeout 300 python bench.py $N1 > gpurun_out/r02_plain32.log 2>&1 &&
timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_tcn|k_ln_stream|k_ln_warp|k_embed|k_pool' -s 34 -c 34 --csv --log-file gpurun_out/r02_launches_stgcn_n32.csv python bench.py $N1 > gpurun_out/r02_ncu32.log 2>&1
# launch list of ONE batch (STGCN_RT_OVERLAP=0): the default two-half step interleaves 45 launches of two CUDA streams
STGCN_RT_OVERLAP=0 timeout 300 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r02_plain_rt.log 2>&1 &&
STGCN_RT_OVERLAP=0 timeout 900 ncu --metrics $M --clock-control none -k regex:'^k_gcnw$|k_rt_|k_embed|k_pool|k_advance' -s 230 -c 23 --csv --log-file gpurun_out/r02_launches_rt4096.csv python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r02_ncu_rt.log 2>&1
timeout 300 python bench.py $N1 > gpurun_out/r02_plain32b.log 2>&1 &&
timeout 1200 ncu --set full --clock-control none -k regex:'^k_gcnw$|k_tcn|k_ln_stream|k_ln_warp' -s 31 -c 31 -o /tmp/r02_full python bench.py $N1 > gpurun_out/r02_ncu_full.log 2>&1 &&
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_fwd32_raw.csv 2>/dev/null
echo done
