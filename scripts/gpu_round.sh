#!/bin/bash
# GPU box: parity tests, the default bench line, the ncu launch list of the same command and
# one --set full capture of the hot kernels at the bench shape (each only after its plain run exited 0).
#   TAG=r2a PYTEST_TARGET=tests/test_gpu_parity.py BENCH_FLAGS="--no-cpu" SKIP_NCU=1 scripts/gpu_round.sh
mkdir -p gpurun_out
T=${TAG:-r2}
if [ -z "$SKIP_TESTS" ]; then
  ( time python -m pytest ${PYTEST_TARGET:-tests/test_gpu_parity.py} -m gpu -x -q --durations=8 ) > gpurun_out/${T}_parity.log 2>&1
  rc=$?; echo "parity rc=$rc"; tail -4 gpurun_out/${T}_parity.log
  [ $rc -ne 0 ] && { grep -n "Error\|assert\|FAILED" gpurun_out/${T}_parity.log | head -30; exit 1; }
fi
python bench.py --steps 3 --warmup 3 ${BENCH_FLAGS---no-cpu} > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
rc=$?; echo "bench rc=$rc"; [ $rc -ne 0 ] && { tail -30 gpurun_out/${T}_bench.err; exit 1; }
python - <<E
import json
d=json.load(open('gpurun_out/${T}_bench.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['roofline']['stage_ms'], d['roofline']['phase_a_ms'], d['roofline']['lookup_stage'])
print(d['roofline']['kernels'])
print(d['e2e']['host_phase_s'])
print('file', d.get('e2e_file'), 'cli', d.get('e2e_cli'), 'pageable', d.get('e2e_pageable'), 'parity', d.get('parity'), 'cpu', d.get('cpu_baseline'))
E
[ -n "$SKIP_NCU" ] && exit 0
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${T}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:${NCU_KERNELS:-sieve_kernel|queue_lookup_kernel|vote_thread_kernel|pack_kernel}" -c ${NCU_COUNT:-4} \
    -o gpurun_out/${T}_prof_resident -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_ncu_a.log 2>&1
echo "ncu resident rc=$?"
ls -la gpurun_out/${T}_*
