#!/bin/bash
# round-2 visit I: the bench on 2 GPUs (torchrun, as the driver launches it) after a 1-GPU run on the same box
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_bench_n1.json 2> gpurun_out/r02i_bench_n1.err; tail -3 gpurun_out/r02i_bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02i_bench_n2.json 2> gpurun_out/r02i_bench_n2.err; tail -3 gpurun_out/r02i_bench_n2.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02i_bench_ref.json 2> gpurun_out/r02i_bench_ref.err; tail -3 gpurun_out/r02i_bench_ref.err
