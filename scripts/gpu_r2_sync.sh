#!/bin/bash
# searches that run together take the filter phase one at a time (BATHHOST_BULK_GATE)
for b in 1 0 1 0; do
  echo "BATHHOST_BULK_GATE=$b"
  BATHHOST_BULK_GATE=$b timeout 300 python scripts/search_concurrent.py 1000 1 2,4,8 2>&1 | tail -1
done
