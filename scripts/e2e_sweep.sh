#!/bin/bash
# GPU box: e2e (utb_search_mem) alone: batch size / slot count / host thread sweeps, and one timeline.
for cfg in "E2E_THREADS=16" "E2E_THREADS=2" "UTB_BATCH_MB=64" "UTB_BATCH_MB=256" "UTB_SLOTS=4" "UTB_SLOTS=8" "UTB_RAMP_MB=16" "E2E_THREADS=2 UTB_HOST_FRAME=1"; do
  echo "== $cfg"
  env $cfg E2E_REPS=4 python scripts/e2e_only.py 2>/dev/null | tail -2
done
echo "== timeline"
UTB_TIMELINE=1 E2E_REPS=2 python scripts/e2e_only.py 2>&1 | tail -24
python scripts/pcie_bw.py
