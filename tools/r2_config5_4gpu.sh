#!/bin/bash
# BASELINE config 5 shape on 4 GPUs (200x pile, -d200 -D5000, 20 % tandem repeats), weak scaling, 3 Mb per GPU (the 500 Mb of the config do not fit the GPU-minute budget)
set -u
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 4 --mb 3 --coverage 200 --depth-cap 200 --maxinput 5000 --repeat-frac 0.2 --steps 2 --warmup 3 --cpu-sample-s 0 --cli 0 --truth-reads 100 2>gpurun_out/cfg5_4gpu.err > gpurun_out/cfg5_4gpu.json
tail -c 1500 gpurun_out/cfg5_4gpu.json; tail -3 gpurun_out/cfg5_4gpu.err
