#!/bin/bash
# GPU box: e2e (utb_search_mem) under different stream-slot counts and batch ramps.
python scripts/pcie_bw.py 2>&1 | tail -3
for cfg in "UTB_SLOTS=6" "UTB_SLOTS=8" "UTB_SLOTS=6 UTB_BATCH_MB=160" "UTB_SLOTS=6 UTB_BATCH_MB=96" "UTB_SLOTS=6 UTB_RAMP_MB=64"; do
  echo "== $cfg"
  env $cfg E2E_REPS=4 python scripts/e2e_only.py 2>/dev/null | tail -2
done
echo "== timeline default"
UTB_TIMELINE=1 E2E_REPS=3 python scripts/e2e_only.py 2>&1 | grep -A 24 "timeline" | tail -22
