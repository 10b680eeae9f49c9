#!/bin/bash
# round-2 closing run: large-target parity, the GPU test suite, the default bench
mkdir -p gpurun_out
timeout 600 python scripts/parity_large.py 300 > gpurun_out/r02w_parity_300mbp.json 2> gpurun_out/r02w_parity.err
tail -1 gpurun_out/r02w_parity_300mbp.json
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02w_bench_n1.json 2> gpurun_out/r02w_bench_n1.err
tail -c 3000 gpurun_out/r02w_bench_n1.json
