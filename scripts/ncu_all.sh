#!/bin/bash
# ncu --set full over every kernel of a small end-to-end search; only the text summary comes back (the report is > 64 MiB)
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:'msv_filter|vit_filter|orf_|fs3_|fs5_|fs_domain|pack_dna4|codon_class|revcomp' -c 44 -f -o /tmp/prof_search \
    python scripts/search_time.py 10 1 > gpurun_out/ncu_search.log 2>&1
tail -3 gpurun_out/ncu_search.log
python scripts/ncu_summary.py /tmp/prof_search.ncu-rep > gpurun_out/search_kernels_ncu.txt
wc -l gpurun_out/search_kernels_ncu.txt
