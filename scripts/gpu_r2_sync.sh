#!/bin/bash
# blocking-sync waits vs spinning in the search leg, packed-upload tests, default bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_packed_upload.py -x -q 2>&1 | tail -2
for b in 1 0; do
  echo "BATHGPU_BLOCKING_SYNC=$b"
  BATHGPU_BLOCKING_SYNC=$b timeout 300 python scripts/search_concurrent.py 1000 1 4,8 2>&1 | tail -1
done
timeout 900 python bench.py > gpurun_out/r02z_bench_n1.json 2> gpurun_out/r02z_bench_n1.err
python -c "
import json; d=json.loads(open('gpurun_out/r02z_bench_n1.json').read().strip().splitlines()[-1]); s=d['search']
print(d['value'], d['e2e']); print({k:s[k] for k in ('value','seconds','one_profile_at_a_time','first_pass_seconds','checks')})"
