#!/bin/bash
# GPU box: rebuild the library with other filter_kernel tunings and time the e2e path (direct filter, 128 MiB batches).
for v in "-DFILT_ILP=2 -DFILT_MINB=4" "-DFILT_ILP=4 -DFILT_MINB=3" "-DFILT_ILP=3 -DFILT_MINB=4"; do
  echo "== $v"
  UTB_NVCC_EXTRA="$v" python -c "from utree_b200 import build; build.build(force=True)" > /dev/null 2>&1
  E2E_REPS=4 python scripts/e2e_only.py 2>/dev/null | tail -2
done
python -c "from utree_b200 import build; build.build(force=True)" > /dev/null 2>&1
