#!/bin/bash
# GPU box: sweep the L2 knobs of the lookup kernel on the L2S workload (2M reads).
for cfg in "" "UTB_L2_PERSIST=0" "UTB_L2_FETCH=32" "UTB_L2_FETCH=128" "UTB_L2_HITRATIO=0.6" "UTB_L2_PERSIST=0 UTB_L2_FETCH=32"; do
  echo "== $cfg"
  env $cfg UTB_STATS=1 python bench.py --steps 3 --warmup 3 --no-cpu --reads 2000000 2> gpurun_out/sweep.err | python -c "
import json,sys;d=json.loads(sys.stdin.read());r=d['roofline'];print(d['value'],r['stage_ms'],'rand32',r['peak'])"
  grep "utree-b200: device" gpurun_out/sweep.err | head -1
done
