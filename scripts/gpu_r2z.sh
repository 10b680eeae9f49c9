#!/bin/bash
# round-2 closing visit: the default bench, then its ncu launch list (final build: concurrent profiles, packed e2e leg)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
tail -c 400 gpurun_out/r02z_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_bench_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 2 > gpurun_out/r02z_ncu_bench.log 2>&1
python scripts/launch_shares.py gpurun_out/r02z_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --search-mbp 2 > gpurun_out/r02z_bench_launch_shares.txt
head -8 gpurun_out/r02z_bench_launch_shares.txt
python -c "
import json; d=json.loads(open('gpurun_out/r02z_bench_n1.json').read().strip().splitlines()[-1]); s=d['search']
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['clocks'], d['gpu_launches']); print({k:s[k] for k in ('value','seconds','one_profile_at_a_time','first_pass_seconds','checks','cpu_baseline')}); print(d['cpu_baseline'])"
