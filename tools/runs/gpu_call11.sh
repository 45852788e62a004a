#!/bin/bash
# round-2 GPU call 11 (2 GPUs): bench.py under torchrun -- trial-sharded headline, T-split and strong-scaling legs
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_c11_bench2.json 2> gpurun_out/r2_c11_bench2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_c11_ref2.json 2> gpurun_out/r2_c11_ref2.err
echo done
