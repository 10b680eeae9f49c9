#!/bin/bash
# round-2 visit K: bench with the integer-filter roofline leg; ncu --set full of the filter / ORF-finder kernels at bench scale (20 Mbp search)
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 10 --warmup 3 --search-mbp 0 --no-cpu-baseline > gpurun_out/r02k_bench_filters.json 2> gpurun_out/r02k_bench_filters.err; tail -5 gpurun_out/r02k_bench_filters.err
ncu --set full --clock-control none -k regex:'msv16_filter|msv_filter|vit_filter|orf_scan|orf_screen|codon_class|bias_forward|orf_forward_parser' -c 12 -f -o /tmp/prof_filters \
    python scripts/search_time.py 20 1 > gpurun_out/r02k_ncu_filters.log 2>&1
tail -3 gpurun_out/r02k_ncu_filters.log
python scripts/ncu_summary.py /tmp/prof_filters.ncu-rep > gpurun_out/r02k_filter_kernels_ncu.txt
wc -l gpurun_out/r02k_filter_kernels_ncu.txt
