#!/bin/bash
# one GPU visit: parity tests, smoke, ncu launch list + full capture of the Forward kernel at the bench workload, then the bench (both arms)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 2 > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fs3_forward -s 3 -c 1 -f -o /tmp/prof_fwd \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 0 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
python scripts/ncu_summary.py /tmp/prof_fwd.ncu-rep > gpurun_out/fwd_ncu_full.txt
python scripts/ncu_source_top.py /tmp/prof_fwd.ncu-rep > gpurun_out/fwd_source_top.txt 2>&1
bash scripts/ncu_all.sh
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>&1; cat gpurun_out/bench_ref.json
