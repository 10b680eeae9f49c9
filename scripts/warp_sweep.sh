for w in 10 11 12 13 14 15 16; do echo "== warps $w"; BATHGPU_FWD_WARPS=$w python scripts/quick_time.py 2>&1 | grep GCUPS | cut -c1-75; done
