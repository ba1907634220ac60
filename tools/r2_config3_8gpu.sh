#!/bin/bash
# BASELINE config 3 shape: one 40x data set, -J sharded over 8 GPUs (strong scaling), 200 Mb in total (1 Gb does not fit the GPU-minute budget)
set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --scaling strong --mb 200 --steps 3 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 0 2>gpurun_out/cfg3_8gpu.err > gpurun_out/cfg3_8gpu.json
tail -c 2500 gpurun_out/cfg3_8gpu.json; tail -3 gpurun_out/cfg3_8gpu.err
