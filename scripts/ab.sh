#!/bin/bash
# A/B timing of Forward-kernel variants: scripts/ab.sh <fwd-version> lib1.so lib2.so ...   (BATHGPU_FWD picks the generation)
ver=$1; shift
for lib in "$@"; do
  echo "=== $lib (BATHGPU_FWD=$ver)"
  BATHGPU_FWD=$ver BATHGPU_LIB=$PWD/$lib python scripts/quick_time.py 2>&1 | grep -v sm_count
done
echo "=== v3 of $1"
BATHGPU_FWD=3 BATHGPU_LIB=$PWD/$1 python scripts/quick_time.py 2>&1 | grep -v sm_count
