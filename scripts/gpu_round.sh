#!/bin/bash
# GPU box: parity tests, the default bench line, and the partition-threshold sweep of the e2e path.
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_parity.py -x -q --durations=5 ) > gpurun_out/r2_parity.log 2>&1
echo "parity rc=$?"; tail -3 gpurun_out/r2_parity.log
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "bench rc=$?"
python - <<'E'
import json
d=json.load(open('gpurun_out/r2_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])
print(d['roofline']['stage_ms'], d['roofline']['phase_a_ms'], d['roofline']['lookup_stage']['survivor_kernel_ms'])
print(d['roofline']['kernels'])
print(d['e2e']['host_phase_s'])
E
for thr in 8 24 48 400; do
  echo "== UTB_PARTITION_MIN_MPOS=$thr"
  UTB_PARTITION_MIN_MPOS=$thr E2E_REPS=3 python scripts/e2e_only.py 2>/dev/null | tail -2
done
