#!/bin/bash
# GPU box: e2e (utb_search_mem) with the newline count on the device (default) or on the host threads,
# with many and with few host threads (E2E_THREADS).
for cfg in "E2E_THREADS=16" "E2E_THREADS=16 UTB_HOST_COUNT=1" "E2E_THREADS=2" "E2E_THREADS=2 UTB_HOST_COUNT=1" "E2E_THREADS=2 UTB_HOST_FRAME=1"; do
  echo "== $cfg"
  env $cfg E2E_REPS=4 python scripts/e2e_only.py 2>/dev/null | tail -2
done
