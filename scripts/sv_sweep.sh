#!/bin/bash
# GPU box: device-resident stage times of the sieve kernel variants (UTB_SV_VARIANT: 0 = U6 at 3 CTAs per SM, 1 = U4/3, 2 = U4/4, 3 = U2/6, 4 = U5/3)
mkdir -p gpurun_out
for v in ${VARIANTS:-0 1 2 3}; do
  UTB_SV_VARIANT=$v python bench.py --steps 3 --warmup 2 --no-cpu --no-extra > gpurun_out/${TAG:-sv}_v$v.json 2> gpurun_out/${TAG:-sv}_v$v.err || { tail -5 gpurun_out/${TAG:-sv}_v$v.err; continue; }
  python - <<E
import json
d=json.load(open('gpurun_out/${TAG:-sv}_v$v.json'))
print('variant $v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['roofline']['stage_ms'], d['roofline']['phase_a_ms'], d['roofline']['lookup_stage']['survivor_kernel_ms'])
E
done
