#!/bin/bash
# GPU box: one bench line per BASELINE config (CONFIGS="l4 long u32"), each with its cpu_baseline; JSON lines
# land in gpurun_out/<TAG>_bench_<config>.json and are copied to profiles/ by hand once they are the final code's.
mkdir -p gpurun_out
T=${TAG:-r2}
for c in ${CONFIGS:-l4 long u32}; do
  python bench.py --config $c --steps ${STEPS:-3} --warmup ${WARMUP:-2} ${BENCH_FLAGS:-} > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err
  rc=$?; echo "bench $c rc=$rc"; [ $rc -ne 0 ] && { tail -25 gpurun_out/${T}_bench_$c.err; continue; }
  python - <<E
import json
d=json.load(open('gpurun_out/${T}_bench_$c.json'))
print('$c', 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'bases/s', d.get('bases_per_s'), 'lookups/s', d['lookups_per_s'])
print('   stage', d['roofline']['stage_ms'], d['roofline']['phase_a_ms'], 'surv', d['roofline']['lookup_stage']['survivor_kernel_ms'], 'hit_rate', d['hit_rate'])
print('   cpu', d.get('cpu_baseline'))
print('   file', d.get('e2e_file'), 'cli', d.get('e2e_cli'), 'pageable', d.get('e2e_pageable'), 'parity', d.get('parity'))
E
done
