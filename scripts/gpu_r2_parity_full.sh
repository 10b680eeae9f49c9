#!/bin/bash
# full-size parity (1 Gbp x 3 profiles, GPU vs CPU backend) and memcheck over the packed-upload kernels
mkdir -p gpurun_out
timeout 900 python scripts/parity_large.py 1000 > gpurun_out/r02z_parity_1gbp.json 2> gpurun_out/r02z_parity_1gbp.err
tail -1 gpurun_out/r02z_parity_1gbp.json | cut -c1-1500
for k in 0 1 2; do [ -f gpurun_out/parity_gpu_$k.tbl ] && diff gpurun_out/parity_gpu_$k.tbl gpurun_out/parity_cpu_$k.tbl > gpurun_out/r02z_parity_1gbp_diff_$k.txt; done
wc -l gpurun_out/r02z_parity_1gbp_diff_*.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_packed_upload.py -x -q -k "not 9_000_001" > gpurun_out/r02z_memcheck_packed.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02z_memcheck_packed.log
tail -3 gpurun_out/r02z_memcheck_packed.log
