#!/bin/bash
# round-2 visit F: device buffers from the stream-ordered pool: all GPU tests, then the search leg with 2 / 4 / 8 contexts per GPU
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for p in 2 4 8; do BATHGPU_TRACE=1 BATHHOST_TRACE=1 timeout 500 python scripts/search_trace.py 1000 $p > gpurun_out/r02f_search_per$p.json 2> gpurun_out/r02f_search_per$p.err; grep -c "device buffer" gpurun_out/r02f_search_per$p.err; done
