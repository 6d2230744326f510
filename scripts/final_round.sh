#!/bin/bash
# GPU box, one GPU: everything profiles/ holds for a round -- the GPU test suite, smoke(), the default bench line with its
# CPU baseline, the reference arm, one line per further BASELINE config, the ncu launch list of the default bench and one
# --set full capture of its hot kernels (both only after the plain run exited 0).  TAG names the outputs in gpurun_out/.
mkdir -p gpurun_out
T=${TAG:-r02}
[ -z "$SKIP_TESTS" ] && { ( time python -m pytest tests -m gpu -q --durations=6 ) > gpurun_out/${T}_gputests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${T}_gputests.log; }
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_l2s.json 2> gpurun_out/${T}_bench_l2s.err; rc=$?; echo "bench rc=$rc"; [ $rc -ne 0 ] && { tail -30 gpurun_out/${T}_bench_l2s.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_bench_reference_arm.err; echo "reference arm rc=$?"
CONFIGS="${CONFIGS:-l4 long u32}" STEPS=3 WARMUP=2 BENCH_FLAGS="" TAG=$T bash scripts/gpu_configs.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/${T}_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:^(pack_kernel|sieve_kernel|queue_lookup_kernel|vote_thread_kernel)" -c 4 \
    -o gpurun_out/${T}_prof_l2s -f python bench.py --steps 1 --warmup 1 --no-cpu --no-extra > gpurun_out/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
python - <<P
import json
d=json.load(open('gpurun_out/${T}_bench_l2s.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d.get('cpu_baseline',{}).get('value'))
print(d['roofline']['kernels'])
for k in ('e2e_file','e2e_cli','e2e_pageable','parity'): print(k, d.get(k))
P
ls -la gpurun_out/${T}_*
