#!/bin/bash
# round-2 kernels under compute-sanitizer: memcheck over the new / rewritten kernels' tests, racecheck over the shared-memory ORF scan
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_orf_finder.py tests/test_gpu_bias_filter.py tests/test_gpu_multidomain_std.py -x -q > gpurun_out/r02r_memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02r_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_orf_finder.py -x -q -k "orfs_and_msv" > gpurun_out/r02r_racecheck.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02r_racecheck.log
tail -4 gpurun_out/r02r_memcheck.log; tail -4 gpurun_out/r02r_racecheck.log
