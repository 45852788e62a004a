#!/bin/bash
# round-2 GPU call 4: where does the fused graph-conv stage lose its time? (measurement build + ncu --set full)
mkdir -p gpurun_out
DBG=$PWD/realtime-st-gcn_b200/csrc/libstgcn_b200_dbg.so
B32="--trials 32 --steps 5 --warmup 3 --no-cpu-baseline --no-rt --no-bf16-leg --no-e2e"
for bits in 0 1024 2048 4096 5120; do
  STGCN_LIB=$DBG STGCN_DEBUG=$bits timeout 300 python bench.py $B32 > gpurun_out/r2_c4_b32_dbg$bits.json 2> gpurun_out/r2_c4_b32_dbg$bits.err
done
timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph > gpurun_out/r2_c4_rt.log 2>&1
STGCN_LIB=$DBG STGCN_DEBUG=8192 timeout 300 python tools/bench_rt.py --streams 4096 --cuda-graph >> gpurun_out/r2_c4_rt.log 2>&1
N1="--trials 32 --steps 1 --warmup 1 --no-rt --no-cpu-baseline --no-bf16-leg --no-e2e"
timeout 300 python bench.py $N1 > gpurun_out/r2_c4_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gcnw<64' -s 3 -c 1 -o gpurun_out/r2_c4_gcnw64 python bench.py $N1 > gpurun_out/r2_c4_ncu.log 2>&1
timeout 300 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c4_plain_rt.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rt_stream<512' -s 60 -c 1 -o gpurun_out/r2_c4_rtstream512 python tools/bench_rt.py --streams 4096 --steps 30 > gpurun_out/r2_c4_ncu_rt.log 2>&1
echo done
