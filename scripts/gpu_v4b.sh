#!/bin/bash
set -x
mkdir -p gpurun_out
BATHGPU_FWD=4 timeout 600 python -m pytest tests/test_gpu_fs_forward.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -3
for v in 4; do BATHGPU_FWD=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --search-mbp 0 > gpurun_out/v4_bench_fwd$v.json 2> gpurun_out/v4_bench_fwd$v.err; tail -2 gpurun_out/v4_bench_fwd$v.err; cat gpurun_out/v4_bench_fwd$v.json; done
BATHGPU_FWD=4 timeout 900 ncu --set full --clock-control none --import-source on -k regex:fs3_forward -s 3 -c 1 -f -o /tmp/prof_fwd4 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 0 > gpurun_out/v4_ncu_full.log 2>&1
tail -3 gpurun_out/v4_ncu_full.log
python scripts/ncu_summary.py /tmp/prof_fwd4.ncu-rep > gpurun_out/v4_fwd_ncu_full.txt
python scripts/ncu_source_top.py /tmp/prof_fwd4.ncu-rep > gpurun_out/v4_fwd_source_top.txt 2>&1
grep -i "issue\|pipe\|stall\|warp cycles\|registers\|occupancy\|eligible\|L1/TEX Hit\|Executed Ipc\|Duration" gpurun_out/v4_fwd_ncu_full.txt | head -60
