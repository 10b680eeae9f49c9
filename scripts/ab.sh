#!/bin/bash
# A/B timing of Forward-kernel variants: scripts/ab.sh lib1.so lib2.so ...   (BATHGPU_FWD picks the generation)
for lib in "$@"; do
  echo "=== $lib"
  BATHGPU_LIB=$PWD/$lib python scripts/quick_time.py 2>&1 | grep -v sm_count
done
echo "=== v1 of $1"
BATHGPU_FWD=1 BATHGPU_LIB=$PWD/$1 python scripts/quick_time.py 2>&1 | grep -v sm_count
