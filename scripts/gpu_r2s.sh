#!/bin/bash
# round-2 visit S: ncu launch list of the bench command and a full capture of the Forward kernel at the bench workload (final build)
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02s_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 2 > gpurun_out/r02s_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fs3_forward -s 3 -c 1 -f -o /tmp/prof_fwd \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 0 --no-filters-leg > gpurun_out/r02s_ncu_full.log 2>&1
tail -3 gpurun_out/r02s_ncu_full.log
python scripts/ncu_summary.py /tmp/prof_fwd.ncu-rep > gpurun_out/r02s_fs3_forward_v3_ncu_full.txt
python scripts/ncu_source_top.py /tmp/prof_fwd.ncu-rep > gpurun_out/r02s_fs3_forward_source_top.txt 2>&1
python scripts/launch_shares.py gpurun_out/r02s_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 2 > gpurun_out/r02s_bench_launch_shares.txt
head -12 gpurun_out/r02s_bench_launch_shares.txt
