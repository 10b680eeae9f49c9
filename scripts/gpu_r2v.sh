#!/bin/bash
# round-2 visit V: final build -- smoke, the bench (both arms) on one GPU
set -x
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 1200 python bench.py > gpurun_out/r02v_bench_n1.json 2> gpurun_out/r02v_bench_n1.err; tail -2 gpurun_out/r02v_bench_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02v_bench_ref.json 2> gpurun_out/r02v_bench_ref.err; tail -2 gpurun_out/r02v_bench_ref.err
