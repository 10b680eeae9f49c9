#!/bin/bash
# round-2 visit P: the bench on 8 GPUs (torchrun, as the driver launches it)
set -x
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02p_bench_n8.json 2> gpurun_out/r02p_bench_n8.err; tail -3 gpurun_out/r02p_bench_n8.err
