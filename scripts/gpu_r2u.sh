#!/bin/bash
# round-2 visit U: multi-warp Backward parser: parity tests, racecheck + memcheck on the long-model cases, the sweep
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_fs_backward.py -x -q -k "synthetic or MET" > gpurun_out/r02u_racecheck_bck_mw.log 2>&1; echo "racecheck rc=$?" | tee -a gpurun_out/r02u_racecheck_bck_mw.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fs_backward.py -x -q -k "synthetic or MET" > gpurun_out/r02u_memcheck_bck_mw.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/r02u_memcheck_bck_mw.log
tail -3 gpurun_out/r02u_racecheck_bck_mw.log; tail -3 gpurun_out/r02u_memcheck_bck_mw.log
timeout 1200 python scripts/gcups_sweep.py > gpurun_out/r02u_gcups_sweep.md 2> gpurun_out/r02u_gcups_sweep.err; grep "| 409\|| 624\|| 903" gpurun_out/r02u_gcups_sweep.md
