#!/bin/bash
# round-2 visit C: all GPU tests (bias filter on the device included), then the bench with the config-4 search leg
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err; tail -5 gpurun_out/r02c_bench_n1.err; cat gpurun_out/r02c_bench_n1.json
BATHHOST_BIAS_HOST=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_n1_biashost.json 2> gpurun_out/r02c_bench_n1_biashost.err
